// Fused equivariant message-passing kernels (the dominant kernels of the hot path).
//
// Reference semantics (as written, per edge e = (j -> i), 32 in-edges per node i):
//   kv   = [onehot(type) | onehot(type) (x) g(|x_i - x_j|) | h_i | h_j]          (340 wide)
//   X2H: k = MLP_k(kv), v = MLP_v(kv) * e_w, q = MLP_q(h)/sqrt(8) (node GEMM)
//        alpha = softmax_j(<q_i, k_ij>) per head ; h_i += sum_j alpha_ij v_ij
//        (repo/modules/attention/x2h_attention.py:43-97)
//   H2X: k = MLP_xk(kv), v = MLP_xv(kv) * e_w  [16 heads]
//        dx_i = mean_heads sum_j alpha_ij v_ij (x_i - x_j)
//        (repo/modules/attention/h2x_attention.py:34-73; x += dx * gen_flag, unitransformer.py:182)
//
// What the kernels do instead (exact algebra, nothing of size [E,*] except alpha*e_w):
//   * first Linear via the node planes: pre = Pi[i] + Pj[j] + c[type] + Wrf[type] g(d)
//   * logits through the query-folded matrix U_i[f][hd] = sum_{f' in hd} q_i[f'] W1k[f'][f]:
//       <q_i, W1k a + b1k>_hd = a . U_i[:,hd] + const(i,hd)     (const cancels in the softmax)
//   * values through linearity of the second Linear:
//       sum_j w_ij (W1v a_ij + b1v) = W1v (sum_j w_ij a_ij) + b1v sum_j w_ij   per head
//   32 in-edges of a node = one warp; in-warp softmax/aggregation, no atomics, fixed
//   summation order (deterministic).
//
// Thread mapping: a warp owns one destination node at a time.  Geometry is computed with
// lane = edge; the MLP math with lane = 4 consecutive features (f = 4*lane .. 4*lane+3) for
// groups of 4 edges; per-(edge, head) contractions are finished with a halving butterfly
// (warp_transpose_reduce) that leaves lane l with edge 4g + l/8 and heads 2*(l%8), 2*(l%8)+1.
#include <math.h>
#include <stdlib.h>
#include "cbg_kernels.cuh"

namespace {

// warps per CTA are a template parameter (8 / 12 / 16): more warps = better latency hiding at the
// price of a tighter register budget (255 / 168 / 128 per thread); picked at run time, see below.
constexpr long long kOffX2hK = cbg_layout::layer_offset(CBG_LF_X2H_K_WRF);
constexpr long long kOffX2hV = cbg_layout::layer_offset(CBG_LF_X2H_V_WRF);
constexpr long long kOffH2x = cbg_layout::layer_offset(CBG_LF_H2X_K_WRF);

struct __align__(16) EdgeMeta {   // per-warp scratch, 3328 B
  float g[CBG_NRBF][32];          // g[m][e]
  float rel[3][32];               // x_i - x_j
  float ew[32];                   // e_w (0 for padded slots)
  int j[32];                      // source node (i itself for padded slots)
  int t[32];                      // edge type 0..3
};

struct MlpSmem {                  // first-layer weights of one edge MLP in shared memory
  const float* wrf;               // [4][20][128]
  const float* c;                 // [4][128]
};

// lane = edge: geometry, RBF, type.  Returns the validity ballot.
__device__ __forceinline__ unsigned edge_setup(EdgeMeta& M, int i, int lane, const float4* __restrict__ x4,
                                               const int* __restrict__ nbr, const float* __restrict__ ew,
                                               const float* s_rbf) {
  const float4 xi = x4[i];
  const int jn = nbr[(size_t)i * CBG_KMAX + lane];
  const bool valid = jn >= 0;
  const int j = valid ? jn : i;
  const float4 xj = x4[j];
  const float rx = xi.x - xj.x, ry = xi.y - xj.y, rz = xi.z - xj.z;
  const float d = sqrtf(rx * rx + ry * ry + rz * rz);
  const float coeff = s_rbf[20];
#pragma unroll
  for (int m = 0; m < CBG_NRBF; ++m) {
    const float u = d - s_rbf[m];
    M.g[m][lane] = expf(coeff * u * u);
  }
  M.rel[0][lane] = rx; M.rel[1][lane] = ry; M.rel[2][lane] = rz;
  M.ew[lane] = valid ? ew[(size_t)i * CBG_KMAX + lane] : 0.f;
  M.j[lane] = j;
  // unitransformer.py:88-99: 0 lig->lig, 1 lig src/prot dst, 2 prot src/lig dst, 3 prot->prot
  M.t[lane] = ((node_flags(xj) & 1) ? 0 : 2) + ((node_flags(xi) & 1) ? 0 : 1);
  const unsigned vmask = __ballot_sync(CBG_FULL, valid);
  __syncwarp();
  return vmask;
}

// First Linear + LayerNorm + ReLU of one edge MLP for the 4 edges e0..e0+3.
// a[ee] = relu(LN(Pi + Pj[j] + c[t] + Wrf[t] g)) restricted to this lane's 4 features.
__device__ __forceinline__ void first_layer4(const EdgeMeta& M, int e0, int lane, const float4 pi,
                                             const float* __restrict__ pj_plane, const MlpSmem W,
                                             const float4 gamma, const float4 beta, float4 (&a)[4]) {
  int t[4];
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) {
    const int j = M.j[e0 + ee];
    t[ee] = M.t[e0 + ee];
    const float4 pj = ldg4(pj_plane + (size_t)j * CBG_H + 4 * lane);
    const float4 c = ld4(W.c + t[ee] * CBG_H + 4 * lane);
    a[ee] = make_float4(pi.x + pj.x + c.x, pi.y + pj.y + c.y, pi.z + pj.z + c.z, pi.w + pj.w + c.w);
  }
  const bool uniform = (t[0] == t[1]) && (t[0] == t[2]) && (t[0] == t[3]);   // warp-uniform
  if (uniform) {
    const float* w = W.wrf + t[0] * (CBG_NRBF * CBG_H) + 4 * lane;
#pragma unroll
    for (int m = 0; m < CBG_NRBF; ++m) {
      const float4 wv = ld4(w + m * CBG_H);
      const float4 gv = ld4(&M.g[m][e0]);
      fma4(a[0], wv, gv.x); fma4(a[1], wv, gv.y); fma4(a[2], wv, gv.z); fma4(a[3], wv, gv.w);
    }
  } else {   // mixed edge types in the group (rare): compact code, per-edge weight rows
#pragma unroll 1
    for (int m = 0; m < CBG_NRBF; ++m) {
      const float4 gv = ld4(&M.g[m][e0]);
      const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
      for (int ee = 0; ee < 4; ++ee)
        fma4(a[ee], ld4(W.wrf + (t[ee] * CBG_NRBF + m) * CBG_H + 4 * lane), gs[ee]);
    }
  }
  // LayerNorm(128, eps=1e-5) over the feature dim (4 per lane x 32 lanes), two-pass
  float s[4];
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) s[ee] = (a[ee].x + a[ee].y) + (a[ee].z + a[ee].w);
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1)
#pragma unroll
    for (int ee = 0; ee < 4; ++ee) s[ee] += __shfl_xor_sync(CBG_FULL, s[ee], m);
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) {
    const float mean = s[ee] * (1.f / 128.f);
    a[ee].x -= mean; a[ee].y -= mean; a[ee].z -= mean; a[ee].w -= mean;
    s[ee] = (a[ee].x * a[ee].x + a[ee].y * a[ee].y) + (a[ee].z * a[ee].z + a[ee].w * a[ee].w);
  }
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1)
#pragma unroll
    for (int ee = 0; ee < 4; ++ee) s[ee] += __shfl_xor_sync(CBG_FULL, s[ee], m);
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) {
    const float rstd = 1.f / sqrtf(s[ee] * (1.f / 128.f) + 1e-5f);
    a[ee].x = fmaxf(fmaf(a[ee].x * rstd, gamma.x, beta.x), 0.f);
    a[ee].y = fmaxf(fmaf(a[ee].y * rstd, gamma.y, beta.y), 0.f);
    a[ee].z = fmaxf(fmaf(a[ee].z * rstd, gamma.z, beta.z), 0.f);
    a[ee].w = fmaxf(fmaf(a[ee].w * rstd, gamma.w, beta.w), 0.f);
  }
}

// U[c][hd] = sum_{d<8} q[hd*8+d] * W1[hd*8+d][4*lane+c]   (W1 natural [f_out][f_in] in smem)
__device__ __forceinline__ void build_u(const float* __restrict__ q_i, const float* s_w1, int lane,
                                        float (&U)[4][CBG_HEADS]) {
#pragma unroll
  for (int hd = 0; hd < CBG_HEADS; ++hd) {
    const float4 q0 = ldg4(q_i + hd * 8), q1 = ldg4(q_i + hd * 8 + 4);
    const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 8; ++d) fma4(u, ld4(s_w1 + (hd * 8 + d) * CBG_H + 4 * lane), qv[d]);
    U[0][hd] = u.x; U[1][hd] = u.y; U[2][hd] = u.z; U[3][hd] = u.w;
  }
}

// part[ee*16+hd] = a[ee] . U[:,hd] (this lane's 4 features), then all-lane sums.
// Afterwards lane l holds, for edge 4g + l/8, heads 2*(l%8) and 2*(l%8)+1 in r0, r1.
__device__ __forceinline__ void contract_heads(const float4 (&a)[4], const float (&U)[4][CBG_HEADS], int lane,
                                               float& r0, float& r1) {
  float part[64];
#pragma unroll
  for (int ee = 0; ee < 4; ++ee)
#pragma unroll
    for (int hd = 0; hd < CBG_HEADS; ++hd)
      part[ee * 16 + hd] = fmaf(a[ee].w, U[3][hd], fmaf(a[ee].z, U[2][hd], fmaf(a[ee].y, U[1][hd], a[ee].x * U[0][hd])));
  warp_transpose_reduce<64>(part, lane);
  r0 = part[0];
  r1 = part[1];
}

// in-warp segment softmax over the 32 edges for this lane's 2 heads; 8 logits per head per lane
// (groups g = 0..7), lanes with equal (lane % 8) share the heads.  Returns alpha in place.
__device__ __forceinline__ void softmax32(float (&lg)[8][2], int lane, unsigned vmask) {
#pragma unroll
  for (int g = 0; g < 8; ++g) {
    const bool valid = (vmask >> (4 * g + (lane >> 3))) & 1u;
    if (!valid) { lg[g][0] = -INFINITY; lg[g][1] = -INFINITY; }
  }
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    float mx = lg[0][i];
#pragma unroll
    for (int g = 1; g < 8; ++g) mx = fmaxf(mx, lg[g][i]);
    mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 8));
    mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 16));
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      lg[g][i] = (mx == -INFINITY) ? 0.f : expf(lg[g][i] - mx);
      sum += lg[g][i];
    }
    sum += __shfl_xor_sync(CBG_FULL, sum, 8);
    sum += __shfl_xor_sync(CBG_FULL, sum, 16);
    const float inv = (sum > 0.f) ? sum : 1.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) lg[g][i] = lg[g][i] / inv;
  }
}

// ------------------------------------------------------------------------------------------------
// X2H, part 1: attention weights  w[i][e][hd] = softmax_e(<q_i, k_ie>) * e_w[i][e]
// smem: K_WRF | K_C | K_LN | K_W1 | K_RBF  (contiguous in the blob) + per-warp EdgeMeta
constexpr int kX2hKFloats = 4 * 20 * 128 + 4 * 128 + 256 + 128 * 128 + 32;
constexpr int x2hk_smem(int w) { return kX2hKFloats * 4 + w * (int)sizeof(EdgeMeta); }

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) x2h_k_kernel(EdgeArgs p) {
  extern __shared__ __align__(16) float smem[];
  const float* s_wrf = smem;
  const float* s_c = s_wrf + 4 * 20 * 128;
  const float* s_ln = s_c + 4 * 128;
  const float* s_w1 = s_ln + 256;
  const float* s_rbf = s_w1 + 128 * 128;
  EdgeMeta* metas = reinterpret_cast<EdgeMeta*>(smem + kX2hKFloats);
  block_copy_f4(smem, p.layer + kOffX2hK, kX2hKFloats);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  EdgeMeta& M = metas[warp];
  const MlpSmem W{s_wrf, s_c};
  const float4 gamma = ld4(s_ln + 4 * lane), beta = ld4(s_ln + 128 + 4 * lane);

  for (int i = blockIdx.x * kWarps + warp; i < p.n_nodes; i += gridDim.x * kWarps) {
    const unsigned vmask = edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf);
    float U[4][CBG_HEADS];
    build_u(p.q + (size_t)i * CBG_H, s_w1, lane, U);
    const float4 pi = ldg4(p.pi_k + (size_t)i * CBG_H + 4 * lane);
    float lg[8][2];
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
      float4 a[4];
      first_layer4(M, 4 * g, lane, pi, p.pj_k, W, gamma, beta, a);
      float r0, r1;
      contract_heads(a, U, lane, r0, r1);
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) if (gg == g) { lg[gg][0] = r0; lg[gg][1] = r1; }
    }
    softmax32(lg, lane, vmask);
    float* wout = p.w + (size_t)i * (CBG_KMAX * CBG_HEADS);
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int e = 4 * g + (lane >> 3);
      const float ew = M.ew[e];
      *reinterpret_cast<float2*>(wout + e * CBG_HEADS + 2 * (lane & 7)) = make_float2(lg[g][0] * ew, lg[g][1] * ew);
    }
    __syncwarp();   // M is rewritten by the next node's setup
  }
}

// ------------------------------------------------------------------------------------------------
// X2H, part 2: h_i += W1v (sum_e w_ie a_ie) + b1v sum_e w_ie   (per head)
// smem: V_WRF | V_C | V_LN | V_W1 | V_B1 | V_RBF + per-warp (EdgeMeta, wbuf[32][16])
constexpr int kX2hVFloats = 4 * 20 * 128 + 4 * 128 + 256 + 128 * 128 + 128 + 32;
constexpr int x2hv_smem(int w) { return kX2hVFloats * 4 + w * ((int)sizeof(EdgeMeta) + 32 * 16 * 4); }

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) x2h_v_kernel(EdgeArgs p) {
  extern __shared__ __align__(16) float smem[];
  const float* s_wrf = smem;
  const float* s_c = s_wrf + 4 * 20 * 128;
  const float* s_ln = s_c + 4 * 128;
  const float* s_w1 = s_ln + 256;
  const float* s_b1 = s_w1 + 128 * 128;
  const float* s_rbf = s_b1 + 128;
  EdgeMeta* metas = reinterpret_cast<EdgeMeta*>(smem + kX2hVFloats);
  float* wbufs = reinterpret_cast<float*>(metas + kWarps);
  block_copy_f4(smem, p.layer + kOffX2hV, kX2hVFloats);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  EdgeMeta& M = metas[warp];
  float* wbuf = wbufs + warp * (32 * 16);
  const MlpSmem W{s_wrf, s_c};
  const float4 gamma = ld4(s_ln + 4 * lane), beta = ld4(s_ln + 128 + 4 * lane);

  for (int i = blockIdx.x * kWarps + warp; i < p.n_nodes; i += gridDim.x * kWarps) {
    edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf);
    {
      const float* wsrc = p.w + (size_t)i * (CBG_KMAX * CBG_HEADS);
#pragma unroll
      for (int r = 0; r < 4; ++r) st4(wbuf + 4 * (lane + 32 * r), ld4(wsrc + 4 * (lane + 32 * r)));
    }
    __syncwarp();
    const float4 pi = ldg4(p.pi_v + (size_t)i * CBG_H + 4 * lane);
    float4 S[CBG_HEADS];
#pragma unroll
    for (int hd = 0; hd < CBG_HEADS; ++hd) S[hd] = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
      float4 a[4];
      first_layer4(M, 4 * g, lane, pi, p.pj_v, W, gamma, beta, a);
#pragma unroll
      for (int ee = 0; ee < 4; ++ee) {
        const float* wr = wbuf + (4 * g + ee) * 16;
        const float4 w0 = ld4(wr), w1 = ld4(wr + 4), w2 = ld4(wr + 8), w3 = ld4(wr + 12);
        const float wv[16] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w,
                              w2.x, w2.y, w2.z, w2.w, w3.x, w3.y, w3.z, w3.w};
#pragma unroll
        for (int hd = 0; hd < CBG_HEADS; ++hd) fma4(S[hd], a[ee], wv[hd]);
      }
    }
    // out[f'] = W1v[f'][:] . S[head(f')][:]  -> 128 partials per lane in two halves of 64
    const float* hin = p.h + (size_t)i * CBG_H;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      float part[64];
#pragma unroll
      for (int fo = 0; fo < 64; ++fo) {
        const int f = half * 64 + fo;
        const float4 wv = ld4(s_w1 + f * CBG_H + 4 * lane);
        const float4 sv = S[f >> 3];
        part[fo] = fmaf(wv.w, sv.w, fmaf(wv.z, sv.z, fmaf(wv.y, sv.y, wv.x * sv.x)));
      }
      warp_transpose_reduce<64>(part, lane);
      const int f0 = half * 64 + 2 * lane;           // this lane's outputs f0, f0+1 (same head)
      const int hd = f0 >> 3;
      float sw = 0.f;
#pragma unroll
      for (int e = 0; e < 32; ++e) sw += wbuf[e * 16 + hd];
      const float2 hv = *reinterpret_cast<const float2*>(hin + f0);
      const float2 b1 = *reinterpret_cast<const float2*>(s_b1 + f0);
      float2 o;
      o.x = hv.x + (part[0] + b1.x * sw);
      o.y = hv.y + (part[1] + b1.y * sw);
      *reinterpret_cast<float2*>(p.h + (size_t)i * CBG_H + f0) = o;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// H2X for the generated nodes only (x moves only where gen_flag, unitransformer.py:182):
//   dx_i = (1/16) sum_hd sum_e alpha_ie^hd e_w (W1xv[hd] . a_v,ie + b1xv[hd]) (x_i - x_j)
// smem: K_WRF|K_C|K_LN|K_W1 | V_WRF|V_C|V_LN|V_W1(16x128)|V_B1(32) | RBF + per-warp EdgeMeta
constexpr int kH2xFloats = (4 * 20 * 128 + 4 * 128 + 256 + 128 * 128) + (4 * 20 * 128 + 4 * 128 + 256 + 16 * 128 + 32) + 32;
constexpr int h2x_smem(int w) { return kH2xFloats * 4 + w * (int)sizeof(EdgeMeta); }

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) h2x_kernel(EdgeArgs p) {
  extern __shared__ __align__(16) float smem[];
  const float* k_wrf = smem;
  const float* k_c = k_wrf + 4 * 20 * 128;
  const float* k_ln = k_c + 4 * 128;
  const float* k_w1 = k_ln + 256;
  const float* v_wrf = k_w1 + 128 * 128;
  const float* v_c = v_wrf + 4 * 20 * 128;
  const float* v_ln = v_c + 4 * 128;
  const float* v_w1 = v_ln + 256;        // [16][128]
  const float* v_b1 = v_w1 + 16 * 128;   // [32]
  const float* s_rbf = v_b1 + 32;
  EdgeMeta* metas = reinterpret_cast<EdgeMeta*>(smem + kH2xFloats);
  block_copy_f4(smem, p.layer + kOffH2x, kH2xFloats);
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  EdgeMeta& M = metas[warp];
  const MlpSmem WK{k_wrf, k_c}, WV{v_wrf, v_c};
  const float4 kga = ld4(k_ln + 4 * lane), kbe = ld4(k_ln + 128 + 4 * lane);
  const float4 vga = ld4(v_ln + 4 * lane), vbe = ld4(v_ln + 128 + 4 * lane);
  const float b1_0 = v_b1[2 * (lane & 7)], b1_1 = v_b1[2 * (lane & 7) + 1];

  for (int n = blockIdx.x * kWarps + warp; n < p.n_nodes; n += gridDim.x * kWarps) {
    const int i = p.node_idx[n];
    const unsigned vmask = edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf);
    float U[4][CBG_HEADS];
    build_u(p.q + (size_t)i * CBG_H, k_w1, lane, U);
    const float4 pik = ldg4(p.pi_k + (size_t)i * CBG_H + 4 * lane);
    const float4 piv = ldg4(p.pi_v + (size_t)i * CBG_H + 4 * lane);
    float lg[8][2], vx[8][2];
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
      float4 a[4];
      first_layer4(M, 4 * g, lane, pik, p.pj_k, WK, kga, kbe, a);
      float r0, r1;
      contract_heads(a, U, lane, r0, r1);
      // dynamic g: keep the register arrays statically indexed
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) if (gg == g) { lg[gg][0] = r0; lg[gg][1] = r1; }
      first_layer4(M, 4 * g, lane, piv, p.pj_v, WV, vga, vbe, a);
      float part[64];
#pragma unroll
      for (int hd = 0; hd < CBG_HEADS; ++hd) {
        const float4 wv = ld4(v_w1 + hd * CBG_H + 4 * lane);
#pragma unroll
        for (int ee = 0; ee < 4; ++ee)
          part[ee * 16 + hd] = fmaf(a[ee].w, wv.w, fmaf(a[ee].z, wv.z, fmaf(a[ee].y, wv.y, a[ee].x * wv.x)));
      }
      warp_transpose_reduce<64>(part, lane);
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) if (gg == g) { vx[gg][0] = part[0] + b1_0; vx[gg][1] = part[1] + b1_1; }
    }
    softmax32(lg, lane, vmask);
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g) {
      const int e = 4 * g + (lane >> 3);
      const float coef = M.ew[e] * (lg[g][0] * vx[g][0] + lg[g][1] * vx[g][1]);
      ax = fmaf(coef, M.rel[0][e], ax);
      ay = fmaf(coef, M.rel[1][e], ay);
      az = fmaf(coef, M.rel[2][e], az);
    }
    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
    if (lane == 0) st4(p.dx + 4 * (size_t)n, make_float4(ax * (1.f / 16.f), ay * (1.f / 16.f), az * (1.f / 16.f), 0.f));
    __syncwarp();
  }
}

int g_num_sms = 0;
int g_edge_warps = 12;

template <int W>
int set_attrs() {
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_k_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, x2hk_smem(W)));
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_v_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, x2hv_smem(W)));
  CBG_CUDA_OK(cudaFuncSetAttribute(h2x_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, h2x_smem(W)));
  return 0;
}

int edge_grid(int n_nodes, int warps) {
  const int need = (n_nodes + warps - 1) / warps;
  return need < g_num_sms ? need : g_num_sms;
}

template <int W>
int launch_x2h(const EdgeArgs& a, cudaStream_t st) {
  const int grid = edge_grid(a.n_nodes, W);
  CBG_PROF_BEGIN(CBG_K_X2H_K, st);
  x2h_k_kernel<W><<<grid, W * 32, x2hk_smem(W), st>>>(a);
  CBG_LAUNCHED(CBG_K_X2H_K, st);
  CBG_PROF_BEGIN(CBG_K_X2H_V, st);
  x2h_v_kernel<W><<<grid, W * 32, x2hv_smem(W), st>>>(a);
  CBG_LAUNCHED(CBG_K_X2H_V, st);
  return 0;
}

template <int W>
int launch_h2x(const EdgeArgs& a, cudaStream_t st) {
  CBG_PROF_BEGIN(CBG_K_H2X, st);
  h2x_kernel<W><<<edge_grid(a.n_nodes, W), W * 32, h2x_smem(W), st>>>(a);
  CBG_LAUNCHED(CBG_K_H2X, st);
  return 0;
}

}  // namespace

int cbg_edge_init(void) {
  static bool done = false;
  if (done) return 0;
  int dev = 0;
  CBG_CUDA_OK(cudaGetDevice(&dev));
  CBG_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  if (const char* e = getenv("CBG_EDGE_WARPS")) {
    const int w = atoi(e);
    if (w == 8 || w == 12 || w == 16) g_edge_warps = w;
  }
  if (int rc = set_attrs<8>()) return rc;
  if (int rc = set_attrs<12>()) return rc;
  if (int rc = set_attrs<16>()) return rc;
  done = true;
  return 0;
}

int cbg_launch_x2h(const EdgeArgs& a, cudaStream_t st) {
  if (a.n_nodes <= 0) return 0;
  if (int rc = cbg_edge_init()) return rc;
  switch (g_edge_warps) {
    case 8: return launch_x2h<8>(a, st);
    case 16: return launch_x2h<16>(a, st);
    default: return launch_x2h<12>(a, st);
  }
}

int cbg_launch_h2x(const EdgeArgs& a, cudaStream_t st) {
  if (a.n_nodes <= 0) return 0;
  if (int rc = cbg_edge_init()) return rc;
  switch (g_edge_warps) {
    case 8: return launch_h2x<8>(a, st);
    case 16: return launch_h2x<16>(a, st);
    default: return launch_h2x<12>(a, st);
  }
}
