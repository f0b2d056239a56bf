"""Index algebra of the tensor-core edge kernels (csrc/edge.cu), checked on the CPU against a numpy model of the
mma.sync.m16n8k8 fragment layouts (PTX ISA: A row-major 16x8: a0 (g,t) a1 (g+8,t) a2 (g,t+4) a3 (g+8,t+4); B col 8x8: b0 (t,g)
b1 (t+4,g); C 16x8: c0 (g,2t) c1 (g,2t+1) c2 (g+8,2t) c3 (g+8,2t+1); g = lane >> 2, t = lane & 3).

The functions below mirror, index for index, what the kernels do: the shared-memory images written by
block_copy_w1k_frag / block_copy_w1v_frag / block_copy_wrf_frag, the per-lane reads, the fragment packing and the
butterfly of the aggregation kernel.  They were written before the first GPU run of those kernels and are kept as a guard
for refactors (the GPU parity tests check the real kernels against the oracle)."""
import numpy as np


def mma(Af, Bf, Cf=None):
    A, B, C = np.zeros((16, 8)), np.zeros((8, 8)), np.zeros((16, 8))
    for l in range(32):
        g, t = l >> 2, l & 3
        A[g, t], A[g + 8, t], A[g, t + 4], A[g + 8, t + 4] = Af[l]
        B[t, g], B[t + 4, g] = Bf[l]
        if Cf is not None:
            C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1] = Cf[l]
    C = C + A @ B
    out = np.zeros((32, 4))
    for l in range(32):
        g, t = l >> 2, l & 3
        out[l] = [C[g, 2 * t], C[g, 2 * t + 1], C[g + 8, 2 * t], C[g + 8, 2 * t + 1]]
    return out


def test_x2h_k_query_folding_and_head_contraction():
    """block_copy_w1k_frag + q staging + U fragments + 64 k-tile MMAs == a . U with U[f][hd] = sum_d q[hd*8+d] W1k[hd*8+d][f]."""
    rng = np.random.default_rng(0)
    W1, q, a = rng.standard_normal((128, 128)), rng.standard_normal(128), rng.standard_normal((32, 128))
    sm = np.full(128 * 128, np.nan)
    for idx in range(128 * 32):                                   # block_copy_w1k_frag
        row, f0 = idx >> 5, (idx & 31) * 4
        hd, d = row >> 3, row & 7
        g, sel = hd & 7, hd >> 3
        base = (g * 8 + d) * 256 + sel
        c0, c1 = (f0 >> 1) ^ (g & 1), ((f0 >> 1) + 1) ^ (g & 1)
        sm[base + c0 * 4], sm[base + c0 * 4 + 2], sm[base + c1 * 4], sm[base + c1 * 4 + 2] = W1[row, f0:f0 + 4]
    assert not np.isnan(sm).any()
    qst = np.full(128, np.nan)
    for lane in range(32):                                        # q staging
        hd, d0 = lane >> 1, 4 * (lane & 1)
        dst = ((hd & 7) * 8 + d0) * 2 + (hd >> 3)
        for k in range(4):
            qst[dst + 2 * k] = q[4 * lane + k]
    Uf = np.zeros((32, 8, 2, 4))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        qp = []
        for dd in range(4):
            v = qst[g * 16 + 4 * dd:g * 16 + 4 * dd + 4]
            qp += [(v[0], v[1]), (v[2], v[3])]
        for d in range(8):
            row = (g * 8 + d) * 256
            for m in range(8):
                for u in range(2):
                    ch = (8 * m + 2 * t + u) ^ (g & 1)
                    w4 = sm[row + 4 * ch:row + 4 * ch + 4]
                    Uf[lane, m, u] += [w4[0] * qp[d][0], w4[1] * qp[d][1], w4[2] * qp[d][0], w4[3] * qp[d][1]]
    acc = np.zeros((4, 32, 4))
    for nt in range(4):
        for m in range(8):
            for u in range(2):
                Bf = np.zeros((32, 2))
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    f = 16 * m + 4 * t + 2 * u
                    Bf[lane] = [a[g + 8 * nt, f], a[g + 8 * nt, f + 1]]
                acc[nt] += mma(Uf[:, m, u], Bf)
    U = np.zeros((128, 16))
    for hd in range(16):
        for d in range(8):
            U[:, hd] += q[hd * 8 + d] * W1[hd * 8 + d]
    want = a @ U
    got = np.zeros((32, 16))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for nt in range(4):
            for hs in range(2):
                for v in range(2):
                    got[8 * nt + 2 * t + v, g + 8 * hs] = acc[nt, lane, 2 * hs + v]
    assert np.abs(got - want).max() < 1e-11


def test_x2h_v_aggregation_and_second_linear():
    """w^T . a as MMAs in the pair layout, then W1v . S from the accumulator fragments against the image of
    block_copy_w1v_frag, reduced over the quad: equals out[f'] = W1v[f'] . S[head(f')]."""
    rng = np.random.default_rng(1)
    W1v, w, a = rng.standard_normal((128, 128)), rng.standard_normal((32, 16)), rng.standard_normal((32, 128))
    sm = np.full(128 * 128, np.nan)
    for idx in range(128 * 32):                                   # block_copy_w1v_frag
        row, f0 = idx >> 5, (idx & 31) * 4
        v, ct = (f0 >> 2) & 1, f0 >> 3
        sw = (row >> 3) & 1
        base = row * 128 + v
        c0, c1 = (2 * ct) ^ sw, (2 * ct + 1) ^ sw
        sm[base + c0 * 4], sm[base + c0 * 4 + 2], sm[base + c1 * 4], sm[base + c1 * 4 + 2] = W1v[row, f0:f0 + 4]
    assert not np.isnan(sm).any()
    acc = np.zeros((4, 4, 32, 4))
    for kt in range(4):
        for c in range(4):
            for qq in range(4):
                Af, Bf = np.zeros((32, 4)), np.zeros((32, 2))
                for lane in range(32):
                    g, t = lane >> 2, lane & 3
                    r0 = 8 * kt + t
                    Af[lane] = [w[r0, g], w[r0, g + 8], w[r0 + 4, g], w[r0 + 4, g + 8]]
                    f = 32 * c + 4 * g + qq
                    Bf[lane] = [a[r0, f], a[r0 + 4, f]]
                acc[c, qq] += mma(Af, Bf)
    part = np.zeros((32, 16))
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        for sel in range(2):
            for d in range(8):
                wrow = ((g + 8 * sel) * 8 + d) * 128
                s = 0.0
                for c in range(4):
                    for hq in range(2):
                        ch = (8 * c + 2 * t + hq) ^ (g & 1)
                        w4 = sm[wrow + 4 * ch:wrow + 4 * ch + 4]
                        s += w4[0] * acc[c, 2 * hq, lane, 2 * sel] + w4[1] * acc[c, 2 * hq, lane, 2 * sel + 1]
                        s += w4[2] * acc[c, 2 * hq + 1, lane, 2 * sel] + w4[3] * acc[c, 2 * hq + 1, lane, 2 * sel + 1]
                part[lane, sel * 8 + d] = s
    xor = lambda arr, mask: np.array([arr[l ^ mask] for l in range(32)])
    p = part.copy()
    for width, mask in ((8, 2), (4, 1)):                           # the two select steps of the quad butterfly
        send, keep = np.zeros((32, width)), np.zeros((32, width))
        for lane in range(32):
            up = (lane & 3 & mask) != 0
            for k in range(width):
                send[lane, k] = p[lane, k] if up else p[lane, k + width]
                keep[lane, k] = p[lane, k + width] if up else p[lane, k]
        for k in range(width):
            p[:, k] = keep[:, k] + xor(send[:, k], mask)
    S = w.T @ a
    want = np.array([W1v[f] @ S[f // 8] for f in range(128)])
    got = np.full(128, np.nan)
    for lane in range(32):
        g, t = lane >> 2, lane & 3
        f0 = (g + 8 * ((t >> 1) & 1)) * 8 + 4 * (t & 1)
        got[f0:f0 + 4] = p[lane, :4]
    assert not np.isnan(got).any() and np.abs(got - want).max() < 1e-11


def test_rbf_matvec_as_masked_mma_accumulates_into_the_activation_fragments():
    """x2h_k_mma2 / h2x_pair: G[16 x 24] . Wrf[type][24 x 128] with rows of other-type / static edges zeroed, once per type,
    through the image of block_copy_wrf_frag, lands in act[m][h] = {(e0,f), (e0,f+1), (e1,f), (e1,f+1)}, f = 16m + 4t + 2h."""
    rng = np.random.default_rng(2)
    Wrf, G = rng.standard_normal((4, 20, 128)), rng.standard_normal((20, 32))
    types, dyn = rng.integers(0, 4, 32), rng.random(32) < 0.7
    img = np.zeros(4 * 3 * 16 * 32 * 2)
    for idx in range(img.size):                                   # block_copy_wrf_frag
        j, ln, nt, kt, ty = idx & 1, (idx >> 1) & 31, (idx >> 6) & 15, (idx >> 10) % 3, idx // (3 << 10)
        gp, tp, m, h = ln >> 2, ln & 3, nt >> 1, nt & 1
        rbf, f = 8 * kt + tp + 4 * j, 16 * m + 4 * (gp >> 1) + 2 * h + (gp & 1)
        img[idx] = Wrf[ty, rbf, f] if rbf < 20 else 0.0
    for pas in range(2):
        act = np.zeros((32, 8, 2, 4))
        for ty in range(4):
            for m in range(8):
                for h in range(2):
                    for kt in range(3):
                        Af, Bf = np.zeros((32, 4)), np.zeros((32, 2))
                        for lane in range(32):
                            g, t = lane >> 2, lane & 3
                            e0, e1 = 16 * pas + g, 16 * pas + g + 8
                            m0 = float(dyn[e0] and types[e0] == ty)
                            m1 = float(dyn[e1] and types[e1] == ty)
                            ra = 8 * kt + t
                            Af[lane] = [G[ra, e0] * m0, G[ra, e1] * m1,
                                        G[ra + 4, e0] * m0 if kt < 2 else 0.0, G[ra + 4, e1] * m1 if kt < 2 else 0.0]
                            base = ty * (3 * 16 * 64) + 2 * lane + (kt * 16 + 2 * m + h) * 64
                            Bf[lane] = img[base:base + 2]
                        act[:, m, h] = mma(Af, Bf, act[:, m, h])
        for lane in range(32):
            g, t = lane >> 2, lane & 3
            e0, e1 = 16 * pas + g, 16 * pas + g + 8
            for m in range(8):
                for h in range(2):
                    for v in range(2):
                        f = 16 * m + 4 * t + 2 * h + v
                        w0 = G[:, e0] @ Wrf[types[e0], :, f] if dyn[e0] else 0.0
                        w1 = G[:, e1] @ Wrf[types[e1], :, f] if dyn[e1] else 0.0
                        assert abs(act[lane, m, h, v] - w0) < 1e-11 and abs(act[lane, m, h, 2 + v] - w1) < 1e-11
