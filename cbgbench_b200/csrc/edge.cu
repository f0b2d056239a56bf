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
// that leaves lane l with head l/2 and edges 4g + 2*(l%2), 4g + 2*(l%2) + 1.  The butterfly is
// select-free for its four head steps because every lane keeps its per-head operand (U, W1xv) in a
// lane-permuted order (physical slot hp <-> head hp ^ (lane >> 1)).
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include "cbg_kernels.cuh"

namespace {

// warps per CTA are a template parameter (8 / 12 / 16): more warps = better latency hiding at the
// price of a tighter register budget (255 / 168 / 128 per thread); picked at run time, see below.
constexpr long long kOffX2hK = cbg_layout::layer_offset(CBG_LF_X2H_K_WRF);
constexpr long long kOffX2hV = cbg_layout::layer_offset(CBG_LF_X2H_V_WRF);
constexpr long long kOffH2x = cbg_layout::layer_offset(CBG_LF_H2X_K_WRF);
constexpr long long kOffX2hKRbf = cbg_layout::layer_offset(CBG_LF_X2H_K_RBF);
constexpr long long kOffX2hVRbf = cbg_layout::layer_offset(CBG_LF_X2H_V_RBF);

struct __align__(16) EdgeMeta {   // per-warp scratch, 3456 B; indexed by the PERMUTED edge position
  float g[CBG_NRBF][32];          // g[m][e]
  float rel[3][32];               // x_i - x_j
  float ew[32];                   // e_w (0 for padded slots)
  int j[32];                      // source node (i itself for padded slots)
  int t[32];                      // edge type 0..3
  int slot[32];                   // index into the node's static-neighbour list (-1: not a static edge)
};

struct MlpSmem {                  // first-layer weights of one edge MLP in shared memory
  const float* wrf;               // [4][20][128]
  const float* c;                 // [4][128]
};

// lane = edge: geometry, RBF, type.  The 32 slots of the node are re-ordered as
//   [static edges (both endpoints have gen_flag == 0) | other valid edges | padding],
// each class keeping its nearest-first order; every edge kernel applies the same permutation, so
// positions agree between x2h_k (writes w) and x2h_v (reads w).  Static edges never change over the
// diffusion steps (their endpoints never move), which is what the R-cache below exploits; the p-th
// static edge of the list is the p-th entry of the node's static-only kNN list (prefix property).
// Returns the validity mask in permuted positions.
// fstat (optional): per-node flag "all 32 slots hold static edges" (written by the edge gate once per step when the
// R-cache is on).  For such a node the permutation is the identity and every edge is served from the R-cache, so only
// j, e_w and the slot index are ever read: the coordinate gathers, the RBF and the type computation are skipped.
__device__ __forceinline__ unsigned edge_setup(EdgeMeta& M, int i, int lane, const float4* __restrict__ x4,
                                               const int* __restrict__ nbr, const float* __restrict__ ew,
                                               const float* s_rbf, const unsigned char* __restrict__ fstat = nullptr) {
  if (fstat != nullptr && fstat[i]) {                     // warp-uniform
    M.j[lane] = nbr[(size_t)i * CBG_KMAX + lane];
    M.ew[lane] = ew[(size_t)i * CBG_KMAX + lane];
    M.slot[lane] = lane;
    __syncwarp();
    return 0xffffffffu;
  }
  const float4 xi = x4[i];
  const int jn = nbr[(size_t)i * CBG_KMAX + lane];
  const bool valid = jn >= 0;
  const int j = valid ? jn : i;
  const float4 xj = x4[j];
  const float rx = xi.x - xj.x, ry = xi.y - xj.y, rz = xi.z - xj.z;
  const float d = sqrtf(rx * rx + ry * ry + rz * rz);
  const int fi = node_flags(xi), fj = node_flags(xj);
  const bool is_static = valid && ((fi | fj) & 2) == 0;
  const unsigned vm = __ballot_sync(CBG_FULL, valid);
  const unsigned sm = __ballot_sync(CBG_FULL, is_static);
  const unsigned lt = (1u << lane) - 1u;
  const int n_static = __popc(sm), n_valid = __popc(vm);
  const int rank_s = __popc(sm & lt);
  const int pos = is_static ? rank_s
                            : (valid ? n_static + __popc(vm & ~sm & lt) : n_valid + __popc(~vm & lt));
  const float coeff = s_rbf[20];
#pragma unroll
  for (int m = 0; m < CBG_NRBF; ++m) {
    const float u = d - s_rbf[m];
    M.g[m][pos] = expf(coeff * u * u);
  }
  M.rel[0][pos] = rx; M.rel[1][pos] = ry; M.rel[2][pos] = rz;
  M.ew[pos] = valid ? ew[(size_t)i * CBG_KMAX + lane] : 0.f;
  M.j[pos] = j;
  // unitransformer.py:88-99: 0 lig->lig, 1 lig src/prot dst, 2 prot src/lig dst, 3 prot->prot
  M.t[pos] = ((fj & 1) ? 0 : 2) + ((fi & 1) ? 0 : 1);
  M.slot[pos] = is_static ? rank_s : -1;
  __syncwarp();
  return (n_valid >= 32) ? 0xffffffffu : ((1u << n_valid) - 1u);
}

// pull the next node's R block (16 KB, contiguous) towards L2 while this node computes
__device__ __forceinline__ void prefetch_rc(const float* rc_base, int i_next, int lane) {
  if (rc_base == nullptr) return;
  const char* b = reinterpret_cast<const char*>(rc_base + (size_t)i_next * (CBG_KMAX * CBG_H));
#pragma unroll
  for (int r = 0; r < 4; ++r) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (size_t)(lane + 32 * r) * 128));
}

// number of entries of the node list of this launch (host bound, optionally clipped by a device count)
__device__ __forceinline__ int list_length(const EdgeArgs& p) {
  int n = p.n_nodes;
  if (p.n_nodes_dev) { const int nd = *p.n_nodes_dev; n = nd < n ? nd : n; }
  return n;
}

// Node scheduling of the X2H kernels.  With a ticket counter (EdgeArgs::ticket, zeroed before the launch) every warp
// draws its next node from a global counter: nodes differ in cost (generated atoms and their neighbourhood run the RBF
// path, static nodes only stream the R-cache), so a static round-robin leaves SMs idle at the end of a launch.  The
// draw for the following node is issued before the current one is processed, which hides the atomic's latency.
struct NodeSched {
  int* ticket;
  int stride, pending, nn;
  __device__ __forceinline__ int first(const EdgeArgs& p, int warp_global, int total_warps, int lane) {
    ticket = p.ticket; stride = total_warps; pending = 0; nn = 0;
    if (ticket == nullptr) return warp_global;
    int v = 0;
    if (lane == 0) v = atomicAdd(ticket, 1);
    return __shfl_sync(CBG_FULL, v, 0);
  }
  __device__ __forceinline__ void draw(int n, int lane) {            // request the node that follows n
    if (ticket != nullptr) { if (lane == 0) pending = atomicAdd(ticket, 1); }
    else pending = n + stride;
  }
  __device__ __forceinline__ int next(int lane) {                    // first use waits for the atomic
    nn = ticket != nullptr ? __shfl_sync(CBG_FULL, pending, 0) : pending;
    return nn;
  }
};

// All-lane sums of 4 per-lane values, result in every lane: transposed butterfly (each step halves
// the number of live values), three plain steps, four broadcasts: 10 SHFL instead of 20.
__device__ __forceinline__ void allreduce4(float (&s)[4], int lane) {
  const bool u1 = (lane & 16) != 0;
  float k0 = u1 ? s[2] : s[0], k1 = u1 ? s[3] : s[1];
  const float d0 = u1 ? s[0] : s[2], d1 = u1 ? s[1] : s[3];
  k0 += __shfl_xor_sync(CBG_FULL, d0, 16);
  k1 += __shfl_xor_sync(CBG_FULL, d1, 16);
  const bool u2 = (lane & 8) != 0;
  float k = u2 ? k1 : k0;
  const float d = u2 ? k0 : k1;
  k += __shfl_xor_sync(CBG_FULL, d, 8);
  k += __shfl_xor_sync(CBG_FULL, k, 4);
  k += __shfl_xor_sync(CBG_FULL, k, 2);
  k += __shfl_xor_sync(CBG_FULL, k, 1);
  s[0] = __shfl_sync(CBG_FULL, k, 0);     // lanes 0-7 hold value 0, 8-15 value 1, 16-23 value 2, 24-31 value 3
  s[1] = __shfl_sync(CBG_FULL, k, 8);
  s[2] = __shfl_sync(CBG_FULL, k, 16);
  s[3] = __shfl_sync(CBG_FULL, k, 24);
}

// First Linear + LayerNorm + ReLU of one edge MLP for the 4 edges e0..e0+3.
// a[ee] = relu(LN(Pi + Pj[j] + c[t] + Wrf[t] g)) restricted to this lane's 4 features.
// rc: this node's block of the R-cache ([32 static slots][128], R = c[t] + Wrf[t] g(d) of the static
// edge) or nullptr.  Groups whose 4 edges are all static skip the RBF mat-vec and stream R instead.
__device__ __forceinline__ void first_layer4(const EdgeMeta& M, int e0, int lane, const float4 pi,
                                             const float* __restrict__ pj_plane, const MlpSmem W,
                                             const float4 gamma, const float4 beta, float4 (&a)[4],
                                             const float* __restrict__ rc) {
  int t[4];
  const bool cached = (rc != nullptr) && (M.slot[e0 + 3] >= 0);   // static edges come first: slot[e0+3]>=0 => all 4
  if (cached) {
#pragma unroll
    for (int ee = 0; ee < 4; ++ee) {
      const int j = M.j[e0 + ee];
      const float4 pj = ldg4(pj_plane + (size_t)j * CBG_H + 4 * lane);
      const float4 r = ldg4(rc + M.slot[e0 + ee] * CBG_H + 4 * lane);
      a[ee] = add4(add4(pi, pj), r);
    }
  } else {
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) {
    const int j = M.j[e0 + ee];
    t[ee] = M.t[e0 + ee];
    const float4 pj = ldg4(pj_plane + (size_t)j * CBG_H + 4 * lane);
    const float4 c = ld4(W.c + t[ee] * CBG_H + 4 * lane);
    a[ee] = add4(add4(pi, pj), c);
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
  }
  // LayerNorm(128, eps=1e-5) over the feature dim (4 per lane x 32 lanes), two-pass
  float s[4];
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) s[ee] = (a[ee].x + a[ee].y) + (a[ee].z + a[ee].w);
  allreduce4(s, lane);
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) {
    a[ee] = add4s(a[ee], -(s[ee] * (1.f / 128.f)));
    s[ee] = dot4(a[ee], a[ee]);
  }
  allreduce4(s, lane);
#pragma unroll
  for (int ee = 0; ee < 4; ++ee) {
    const float rstd = 1.f / sqrtf(s[ee] * (1.f / 128.f) + 1e-5f);
    a[ee] = ln_relu4(a[ee], rstd, gamma, beta);
  }
}

// Query-folded key matrix, lane-permuted: Up[c][hp] = U[c][hp ^ (lane>>1)] with
// U[c][hd] = sum_{d<8} q[hd*8+d] * W1[hd*8+d][4*lane+c]   (W1 natural [f_out][f_in] in smem).
__device__ __forceinline__ void build_u(const float* __restrict__ q_i, const float* s_w1, int lane,
                                        float (&U)[4][CBG_HEADS]) {
  const int hx = lane >> 1;
#pragma unroll
  for (int hp = 0; hp < CBG_HEADS; ++hp) {
    const int hd = hp ^ hx;
    const float4 q0 = ldg4(q_i + hd * 8), q1 = ldg4(q_i + hd * 8 + 4);
    const float qv[8] = {q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, q1.z, q1.w};
    float4 u = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int d = 0; d < 8; ++d) fma4(u, ld4(s_w1 + (hd * 8 + d) * CBG_H + 4 * lane), qv[d]);
    U[0][hp] = u.x; U[1][hp] = u.y; U[2][hp] = u.z; U[3][hp] = u.w;
  }
}

// Streaming head contraction + warp reduction.  For the 4 edges of a group and the 16 lane-permuted head
// slots hp, leaf(hp)[ee] = a[ee] . W(hp) (this lane's 4 features) must be summed over all lanes.  The
// halving butterfly  level 0: hp vs hp+8 (xor 16), 1: hp+4 (xor 8), 2: hp+2 (xor 4), 3: hp+1 (xor 2)  is
// evaluated depth first, so a partial result is shuffled as soon as its two halves exist: only ~24
// values are live instead of 64 and the shuffles overlap the FMAs of the next leaves.  Because slot hp
// holds head hp ^ (lane>>1), every lane keeps its lower half and sends its upper half: no selects.
template <int LVL, int HP, class Leaf>
struct HeadReduce {
  static __device__ __forceinline__ void run(const Leaf& leaf, float (&out)[4]) {
    float lo[4], hi[4];
    HeadReduce<LVL + 1, HP, Leaf>::run(leaf, lo);
    HeadReduce<LVL + 1, HP + (8 >> LVL), Leaf>::run(leaf, hi);
#pragma unroll
    for (int ee = 0; ee < 4; ++ee) out[ee] = lo[ee] + __shfl_xor_sync(CBG_FULL, hi[ee], 16 >> LVL);
  }
};
// level 3 evaluates its two leaves (head slots HP, HP+1) together so the dot products can use packed FMAs
template <int HP, class Leaf>
struct HeadReduce<3, HP, Leaf> {
  static __device__ __forceinline__ void run(const Leaf& leaf, float (&out)[4]) {
    float lo[4], hi[4];
    leaf.template eval2<HP>(lo, hi);
#pragma unroll
    for (int ee = 0; ee < 4; ++ee) out[ee] = lo[ee] + __shfl_xor_sync(CBG_FULL, hi[ee], 2);
  }
};
// last step: split on the high edge bit (lane bit 0).  r0, r1 = totals for head lane>>1 and edges
// 2*(lane&1), 2*(lane&1)+1 of the group.
template <class Leaf>
__device__ __forceinline__ void reduce_heads(const Leaf& leaf, int lane, float& r0, float& r1) {
  float f[4];
  HeadReduce<0, 0, Leaf>::run(leaf, f);
  const bool up = (lane & 1) != 0;
  const float s0 = up ? f[0] : f[2], s1 = up ? f[1] : f[3];
  const float k0 = up ? f[2] : f[0], k1 = up ? f[3] : f[1];
  r0 = k0 + __shfl_xor_sync(CBG_FULL, s0, 1);
  r1 = k1 + __shfl_xor_sync(CBG_FULL, s1, 1);
}

// leaf = a[ee] . Up[:,hp] with the lane-permuted, query-folded key matrix in registers
struct ULeaf {
  const float4 (&a)[4];
  const float (&U)[4][CBG_HEADS];
  template <int HP>
  __device__ __forceinline__ void eval2(float (&lo)[4], float (&hi)[4]) const {
#pragma unroll
    for (int ee = 0; ee < 4; ++ee) {
      float2 t = __fmul2_rn(make_float2(U[0][HP], U[0][HP + 1]), make_float2(a[ee].x, a[ee].x));
      t = __ffma2_rn(make_float2(U[1][HP], U[1][HP + 1]), make_float2(a[ee].y, a[ee].y), t);
      t = __ffma2_rn(make_float2(U[2][HP], U[2][HP + 1]), make_float2(a[ee].z, a[ee].z), t);
      t = __ffma2_rn(make_float2(U[3][HP], U[3][HP + 1]), make_float2(a[ee].w, a[ee].w), t);
      lo[ee] = t.x;
      hi[ee] = t.y;
    }
  }
};
// leaf = a[ee] . W[head][:] with the weight rows in shared memory ([16][128], lane-permuted head slot)
struct SmemLeaf {
  const float4 (&a)[4];
  const float* w;      // base of the [16][128] matrix
  int lane;
  template <int HP>
  __device__ __forceinline__ void eval2(float (&lo)[4], float (&hi)[4]) const {
    const float4 w0 = ld4(w + (HP ^ (lane >> 1)) * CBG_H + 4 * lane);
    const float4 w1 = ld4(w + ((HP + 1) ^ (lane >> 1)) * CBG_H + 4 * lane);
#pragma unroll
    for (int ee = 0; ee < 4; ++ee) { lo[ee] = dot4(a[ee], w0); hi[ee] = dot4(a[ee], w1); }
  }
};

__device__ __forceinline__ void contract_heads(const float4 (&a)[4], const float (&U)[4][CBG_HEADS], int lane,
                                               float& r0, float& r1) {
  reduce_heads(ULeaf{a, U}, lane, r0, r1);
}

// edge handled by this lane for value i (0/1) of group g under the reduce_heads mapping
__device__ __forceinline__ int lane_edge(int g, int lane, int i) { return 4 * g + 2 * (lane & 1) + i; }

// in-warp segment softmax over the 32 edges of head lane>>1: 16 logits per lane (8 groups x 2
// edges), the other 16 live in lane^1.  Returns alpha in place.
__device__ __forceinline__ void softmax32(float (&lg)[8][2], int lane, unsigned vmask) {
  float mx = -INFINITY;
#pragma unroll
  for (int g = 0; g < 8; ++g)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const bool valid = (vmask >> lane_edge(g, lane, i)) & 1u;
      if (!valid) lg[g][i] = -INFINITY;
      mx = fmaxf(mx, lg[g][i]);
    }
  mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 1));
  float sum = 0.f;
#pragma unroll
  for (int g = 0; g < 8; ++g)
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      lg[g][i] = (mx == -INFINITY) ? 0.f : expf(lg[g][i] - mx);
      sum += lg[g][i];
    }
  sum += __shfl_xor_sync(CBG_FULL, sum, 1);
  const float inv = (sum > 0.f) ? sum : 1.f;
#pragma unroll
  for (int g = 0; g < 8; ++g)
#pragma unroll
    for (int i = 0; i < 2; ++i) lg[g][i] = lg[g][i] / inv;
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

  const int n_list = list_length(p);
  NodeSched sch;
  for (int n = sch.first(p, blockIdx.x * kWarps + warp, gridDim.x * kWarps, lane); n < n_list; n = sch.nn) {
    sch.draw(n, lane);
    const int i = p.node_idx ? p.node_idx[n] : n;
    const unsigned vmask = edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf, p.fstat);
    const float* rc = p.rc_k ? p.rc_k + (size_t)i * (CBG_KMAX * CBG_H) : nullptr;
    {
      const int nn = sch.next(lane);
      if (nn < n_list) prefetch_rc(p.rc_k, p.node_idx ? p.node_idx[nn] : nn, lane);
    }
    float U[4][CBG_HEADS];
    build_u(p.q + (size_t)i * CBG_H, s_w1, lane, U);
    const float4 pi = ldg4(p.pi_k + (size_t)i * CBG_H + 4 * lane);
    float lg[8][2];
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
      float4 a[4];
      first_layer4(M, 4 * g, lane, pi, p.pj_k, W, gamma, beta, a, rc);
      float r0, r1;
      contract_heads(a, U, lane, r0, r1);
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) if (gg == g) { lg[gg][0] = r0; lg[gg][1] = r1; }
    }
    softmax32(lg, lane, vmask);
    float* wout = p.w + (size_t)i * (CBG_KMAX * CBG_HEADS);
#pragma unroll
    for (int g = 0; g < 8; ++g)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = lane_edge(g, lane, i);
        wout[e * CBG_HEADS + (lane >> 1)] = lg[g][i] * M.ew[e];
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

  const int n_list = list_length(p);
  NodeSched sch;
  for (int n = sch.first(p, blockIdx.x * kWarps + warp, gridDim.x * kWarps, lane); n < n_list; n = sch.nn) {
    sch.draw(n, lane);
    const int i = p.node_idx ? p.node_idx[n] : n;
    edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf, p.fstat);
    const float* rc = p.rc_v ? p.rc_v + (size_t)i * (CBG_KMAX * CBG_H) : nullptr;
    {
      const int nn = sch.next(lane);
      if (nn < n_list) prefetch_rc(p.rc_v, p.node_idx ? p.node_idx[nn] : nn, lane);
    }
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
      first_layer4(M, 4 * g, lane, pi, p.pj_v, W, gamma, beta, a, rc);
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
    // out[f'] = W1v[f'][:] . S[head(f')][:]  -> 128 partials per lane in two halves of 64.
    // Partial index = dp*8 + hh with f' = half*64 + hh*8 + (dp ^ (lane>>2)): the three steps over the
    // within-head bits are select-free (lane-permuted rows of W1v), the two steps over head bits
    // use selects.  Lane l ends with hh = 2*(l&3) + {0,1}, d = l>>2.
    const float* hin = p.h + (size_t)i * CBG_H;
    const int dx3 = lane >> 2;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      // g[hh] = sum over lanes of W1v[f(dp,hh)] . S[half*8+hh], reduced over the three within-head bits
      // (dp vs dp+4: xor 16, dp+2: xor 8, dp+1: xor 4) depth first, then the two head-bit steps with selects
      float part[8];
#pragma unroll
      for (int hh = 0; hh < 8; ++hh) {
        const float4 sv = S[half * 8 + hh];
        float lv[8];
#pragma unroll
        for (int dp = 0; dp < 8; ++dp) {
          const int f = half * 64 + hh * 8 + (dp ^ dx3);
          const float4 wv = ld4(s_w1 + f * CBG_H + 4 * lane);
          lv[dp] = dot4(wv, sv);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) lv[k] += __shfl_xor_sync(CBG_FULL, lv[k + 4], 16);
#pragma unroll
        for (int k = 0; k < 2; ++k) lv[k] += __shfl_xor_sync(CBG_FULL, lv[k + 2], 8);
        part[hh] = lv[0] + __shfl_xor_sync(CBG_FULL, lv[1], 4);
      }
      {   // hh bit 2 <-> lane bit 1
        const bool up = (lane & 2) != 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float send = up ? part[k] : part[k + 4];
          const float keep = up ? part[k + 4] : part[k];
          part[k] = keep + __shfl_xor_sync(CBG_FULL, send, 2);
        }
      }
      {   // hh bit 1 <-> lane bit 0
        const bool up = (lane & 1) != 0;
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float send = up ? part[k] : part[k + 2];
          const float keep = up ? part[k + 2] : part[k];
          part[k] = keep + __shfl_xor_sync(CBG_FULL, send, 1);
        }
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int hh = 2 * (lane & 3) + k;
        const int hd = half * 8 + hh;
        const int f = half * 64 + hh * 8 + dx3;
        float sw = 0.f;
#pragma unroll
        for (int e = 0; e < 32; ++e) sw += wbuf[e * 16 + hd];
        p.h[(size_t)i * CBG_H + f] = hin[f] + (part[k] + s_b1[f] * sw);
      }
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
  const float b1 = v_b1[lane >> 1];      // this lane's head under the reduce_heads mapping

  for (int n = blockIdx.x * kWarps + warp; n < p.n_nodes; n += gridDim.x * kWarps) {
    const int i = p.node_idx[n];
    const unsigned vmask = edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf, p.fstat);
    float U[4][CBG_HEADS];
    build_u(p.q + (size_t)i * CBG_H, k_w1, lane, U);
    const float4 pik = ldg4(p.pi_k + (size_t)i * CBG_H + 4 * lane);
    const float4 piv = ldg4(p.pi_v + (size_t)i * CBG_H + 4 * lane);
    float lg[8][2], vx[8][2];
#pragma unroll 1
    for (int g = 0; g < 8; ++g) {
      float4 a[4];
      first_layer4(M, 4 * g, lane, pik, p.pj_k, WK, kga, kbe, a, nullptr);
      float r0, r1;
      contract_heads(a, U, lane, r0, r1);
      // dynamic g: keep the register arrays statically indexed
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) if (gg == g) { lg[gg][0] = r0; lg[gg][1] = r1; }
      first_layer4(M, 4 * g, lane, piv, p.pj_v, WV, vga, vbe, a, nullptr);
      reduce_heads(SmemLeaf{a, v_w1, lane}, lane, r0, r1);
#pragma unroll
      for (int gg = 0; gg < 8; ++gg) if (gg == g) { vx[gg][0] = r0 + b1; vx[gg][1] = r1 + b1; }
    }
    softmax32(lg, lane, vmask);
    float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
    for (int g = 0; g < 8; ++g)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int e = lane_edge(g, lane, i);
        const float coef = M.ew[e] * lg[g][i] * vx[g][i];
        ax = fmaf(coef, M.rel[0][e], ax);
        ay = fmaf(coef, M.rel[1][e], ay);
        az = fmaf(coef, M.rel[2][e], az);
      }
    ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
    if (lane == 0) st4(p.dx + 4 * (size_t)n, make_float4(ax * (1.f / 16.f), ay * (1.f / 16.f), az * (1.f / 16.f), 0.f));
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// R-cache build (once per batch): R[i][s][:] = c[t] + Wrf[t] g(|x_i - x_j|) for the s-th entry j of
// node i's static-only neighbour list.  blockIdx.y selects (layer, k|v MLP).
constexpr int kRcFloats = 4 * 20 * 128 + 4 * 128;   // Wrf | c (contiguous in the blob)
constexpr int kRcSmem = (kRcFloats + 32) * 4 + 8 * (20 * 32 + 32) * 4;

__global__ void __launch_bounds__(256) rcache_kernel(const float* __restrict__ layers, long long layer_stride,
                                                     const float4* __restrict__ x4, const int* __restrict__ snbr,
                                                     int n_nodes, float* __restrict__ rcache) {
  extern __shared__ __align__(16) float smem[];
  const int which = blockIdx.y & 1, layer = blockIdx.y >> 1;
  const float* L = layers + (size_t)layer * layer_stride;
  const float* wsrc = L + (which ? kOffX2hV : kOffX2hK);
  const float* rbf_src = L + (which ? kOffX2hVRbf : kOffX2hKRbf);
  float* s_wrf = smem;
  float* s_c = smem + 4 * 20 * 128;
  float* s_rbf = smem + kRcFloats;
  block_copy_f4(smem, wsrc, kRcFloats);
  if (threadIdx.x < 32) s_rbf[threadIdx.x] = rbf_src[threadIdx.x];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float* g = smem + kRcFloats + 32 + warp * (20 * 32 + 32);   // g[m][s]
  int* ts = reinterpret_cast<int*>(g + 20 * 32);
  float* out_base = rcache + (size_t)blockIdx.y * n_nodes * (CBG_KMAX * CBG_H);
  for (int i = blockIdx.x * 8 + warp; i < n_nodes; i += gridDim.x * 8) {
    const float4 xi = x4[i];
    if (node_flags(xi) & 2) continue;                   // moving centre: no static edges
    const int jn = snbr[(size_t)i * CBG_KMAX + lane];
    const int j = jn >= 0 ? jn : i;
    const float4 xj = x4[j];
    const float rx = xi.x - xj.x, ry = xi.y - xj.y, rz = xi.z - xj.z;
    const float d = sqrtf(rx * rx + ry * ry + rz * rz);
    const float coeff = s_rbf[20];
#pragma unroll
    for (int m = 0; m < CBG_NRBF; ++m) { const float u = d - s_rbf[m]; g[m * 32 + lane] = expf(coeff * u * u); }
    ts[lane] = ((node_flags(xj) & 1) ? 0 : 2) + ((node_flags(xi) & 1) ? 0 : 1);
    __syncwarp();
    float* out = out_base + (size_t)i * (CBG_KMAX * CBG_H);
    for (int s0 = 0; s0 < 32; s0 += 4) {
      float4 a[4];
#pragma unroll
      for (int ee = 0; ee < 4; ++ee) a[ee] = ld4(s_c + ts[s0 + ee] * CBG_H + 4 * lane);
#pragma unroll 4
      for (int m = 0; m < CBG_NRBF; ++m) {
        const float4 gv = ld4(g + m * 32 + s0);
        const float gs[4] = {gv.x, gv.y, gv.z, gv.w};
#pragma unroll
        for (int ee = 0; ee < 4; ++ee)
          fma4(a[ee], ld4(s_wrf + (ts[s0 + ee] * CBG_NRBF + m) * CBG_H + 4 * lane), gs[ee]);
      }
#pragma unroll
      for (int ee = 0; ee < 4; ++ee) st4(out + (s0 + ee) * CBG_H + 4 * lane, a[ee]);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// Tensor-core variants of the two X2H kernels.
//
// The per-edge contractions of X2H are small GEMMs per destination node:
//   x2h_k:  logits^T[16 heads][32 edges] = U_i^T[16][128] . a^T[128][32]     (M=16, N=32, K=128)
//   x2h_v:  S[16 heads][128 features]    = w_i^T[16][32]  . a[32][128]       (M=16, N=128, K=32)
// with a[e][:] = relu(LN(first Linear)) of edge e.  They run on the tensor cores with
// mma.sync.m16n8k8 TF32 and the 3xTF32 split (a_lo*b_hi + a_hi*b_lo + a_hi*b_hi, fp32 accumulate), which keeps
// fp32-level accuracy.  The operands differ per node (U_i, w_i), so there is nothing for tcgen05 / TMA to share
// across a 128-row tile; a warp-level MMA per node is the matching granularity.
//
// What changes against the SIMT kernels above is the thread mapping: the MMA fragment layout decides which lane
// owns which (edge, feature) pairs, so the whole front end (gather, first Linear, LayerNorm) is computed directly in
// fragment layout and `a` never leaves registers:
//   x2h_k: lane (g = lane>>2, t = lane&3) owns edges g + 8*nt (nt = 0..3) and the 32 features 16m + 4t + q;
//          the LayerNorm statistics of an edge are a quad reduction (2 shuffles)
//   x2h_v: lane (g, t) owns edges 8*kt + t and 8*kt + t + 4 (kt = 0..3) and the 16 features 32c + 4g + q;
//          the statistics are a reduction over the 8 lanes with the same t (3 shuffles)
// Edges are processed in their permuted positions (edge_setup); a block of 8 consecutive positions is served from
// the R-cache when all 8 are static (warp-uniform test), else its RBF mat-vec runs in registers.
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], const unsigned (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
// v = hi + lo exactly, hi = the TF32 the tensor core reads (top 19 bits); lo is read truncated (error 2^-21 |v|)
__device__ __forceinline__ void split_tf32(const float v, unsigned& hi, unsigned& lo) {
  hi = __float_as_uint(v) & 0xffffe000u;
  lo = __float_as_uint(v - __uint_as_float(hi));
}
// packed fp32 pairs as opaque 64-bit registers (PTX f32x2): keeps a pair packed across many uses
__device__ __forceinline__ unsigned long long pack_f32x2(const float lo, const float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack_f32x2(const unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ unsigned long long fma_f32x2(const unsigned long long a, const unsigned long long b,
                                                        const unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
template <int N>
__device__ __forceinline__ void split_frag(const float (&v)[N], unsigned (&hi)[N], unsigned (&lo)[N]) {
#pragma unroll
  for (int k = 0; k < N; ++k) split_tf32(v[k], hi[k], lo[k]);
}
// c += A.B to fp32 accuracy: the two cross terms first, the leading term last
__device__ __forceinline__ void mma3(float (&c)[4], const unsigned (&ah)[4], const unsigned (&al)[4],
                                     const unsigned (&bh)[2], const unsigned (&bl)[2]) {
  mma_tf32(c, al, bh);
  mma_tf32(c, ah, bl);
  mma_tf32(c, ah, bh);
}
__device__ __forceinline__ float comp4(const float4 v, const int k) {   // k is a compile-time constant after unrolling
  return k == 0 ? v.x : (k == 1 ? v.y : (k == 2 ? v.z : v.w));
}

// Shared-memory images of the two [128][128] second-Linear matrices, laid out so that every operand the kernels
// need is one LDS.128 that lands in consecutive registers (no register shuffling before an FFMA2 / MMA) and so that
// the 8 lanes of an LDS.128 phase (two heads g, g+1 x four t) hit disjoint banks (16-byte chunk index XOR (g & 1)).
//
// x2h_k: W1k[hd*8+d][f] -> the (g = hd & 7, d) row holds 128 x (f, sel = hd >> 3) interleaved: element
//   ((g*8+d)*64 + ((f >> 1) ^ (g & 1)))*4 + (f & 1)*2 + sel, so one chunk = {(f,hd g), (f,hd g+8), (f+1,hd g), (f+1,hd g+8)}
//   = the A fragment (a0..a3) of the k-tile whose slots t / t+4 are the features f / f+1.
__device__ __forceinline__ void block_copy_w1k_frag(float* dst, const float* __restrict__ src) {
  for (int idx = threadIdx.x; idx < 128 * 32; idx += blockDim.x) {
    const int row = idx >> 5, f0 = (idx & 31) * 4;
    const int hd = row >> 3, d = row & 7, g = hd & 7, sel = hd >> 3;
    const float4 v = ldg4(src + 4 * idx);
    float* base = dst + (g * 8 + d) * 256 + sel;
    const int c0 = (f0 >> 1) ^ (g & 1), c1 = ((f0 >> 1) + 1) ^ (g & 1);
    base[c0 * 4] = v.x; base[c0 * 4 + 2] = v.y; base[c1 * 4] = v.z; base[c1 * 4 + 2] = v.w;
  }
}
// x2h_v: W1v[row][f], f = 32c + 8t + 4v + q -> element row*128 + 4*((8c + 2t + (q >> 1)) ^ ((row >> 3) & 1)) + 2*(q & 1) + v,
//   so one chunk = {(q,v=0), (q,v=1), (q+1,v=0), (q+1,v=1)}: pairs over v, matching the accumulator pairs (c0,c1)/(c2,c3).
__device__ __forceinline__ void block_copy_w1v_frag(float* dst, const float* __restrict__ src) {
  for (int idx = threadIdx.x; idx < 128 * 32; idx += blockDim.x) {
    const int row = idx >> 5, f0 = (idx & 31) * 4;       // f0 = 32c + 8t + 4v, q = 0..3
    const int v = (f0 >> 2) & 1, ct = f0 >> 3;           // ct = 4c + t
    const float4 x = ldg4(src + 4 * idx);
    const int sw = (row >> 3) & 1;
    float* base = dst + row * CBG_H + v;
    const int c0 = (2 * ct) ^ sw, c1 = (2 * ct + 1) ^ sw;
    base[c0 * 4] = x.x; base[c0 * 4 + 2] = x.y; base[c1 * 4] = x.z; base[c1 * 4 + 2] = x.w;
  }
}

constexpr int x2hk_mma_smem(int w) { return kX2hKFloats * 4 + w * ((int)sizeof(EdgeMeta) + 256 * 4); }   // + q_i | Pi_i staging
constexpr int x2hv_mma_smem(int w) { return kX2hVFloats * 4 + w * ((int)sizeof(EdgeMeta) + 128 * 4); }   // + Pi_i staging

// RP = edge blocks (rows of this lane) per pass: 2 keeps 64 registers of activations live and splits U once per two
// rows; 1 halves that (fits 12 warps per CTA) and splits U once per row.
template <int kWarps, int RP>
__global__ void __launch_bounds__(kWarps * 32, 1) x2h_k_mma_kernel(EdgeArgs p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int kHead = 4 * 20 * 128 + 4 * 128 + 256;     // WRF | C | LN
  const float* s_wrf = smem;
  const float* s_c = s_wrf + 4 * 20 * 128;
  const float* s_ln = s_c + 4 * 128;
  const float* s_w1 = s_ln + 256;                          // fragment layout, see block_copy_w1*_frag
  const float* s_rbf = s_w1 + 128 * 128;
  EdgeMeta* metas = reinterpret_cast<EdgeMeta*>(smem + kX2hKFloats);
  {
    const float* src = p.layer + kOffX2hK;
    block_copy_f4(smem, src, kHead);
    block_copy_w1k_frag(smem + kHead, src + kHead);
    block_copy_f4(smem + kHead + 128 * 128, src + kHead + 128 * 128, 32);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  EdgeMeta& M = metas[warp];
  float* qst = reinterpret_cast<float*>(metas + kWarps) + warp * 256;   // per-warp staging of q_i (fragment order) | Pi_i
  float* pist = qst + 128;

  const int n_list = list_length(p);
  NodeSched sch;
  for (int n = sch.first(p, blockIdx.x * kWarps + warp, gridDim.x * kWarps, lane); n < n_list; n = sch.nn) {
    sch.draw(n, lane);
    const int i = p.node_idx ? p.node_idx[n] : n;
    const unsigned vmask = edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf, p.fstat);
    const float* rc = p.rc_k ? p.rc_k + (size_t)i * (CBG_KMAX * CBG_H) : nullptr;
    const int nst = rc ? __popc(__ballot_sync(CBG_FULL, M.slot[lane] >= 0)) : 0;
    {
      const int nn = sch.next(lane);
      if (nn < n_list) prefetch_rc(p.rc_k, p.node_idx ? p.node_idx[nn] : nn, lane);
    }
    // query-folded key matrix directly as A fragments: Uf[m][u] = {U[f][g], U[f][g+8], U[f+1][g], U[f+1][g+8]},
    // f = 16m + 4t + 2u, U[f][hd] = sum_d q[hd*8+d] W1k[hd*8+d][f]
    float4 Uf[8][2];
    {
      // stage q_i through shared memory so that (q of head g, q of head g+8) arrive as adjacent register pairs:
      // element hd*8+d -> ((hd & 7)*8 + d)*2 + (hd >> 3)
      {
        const float4 q4 = ldg4(p.q + (size_t)i * CBG_H + 4 * lane);
        const int hd = lane >> 1, d0 = 4 * (lane & 1);
        float* dst = qst + ((hd & 7) * 8 + d0) * 2 + (hd >> 3);
        dst[0] = q4.x; dst[2] = q4.y; dst[4] = q4.z; dst[6] = q4.w;
        st4(pist + 4 * lane, ldg4(p.pi_k + (size_t)i * CBG_H + 4 * lane));   // Pi row: read back per edge block via LDS
      }
      __syncwarp();
      unsigned long long qp[8];
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {
        const float4 v = ld4(qst + g * 16 + 4 * dd);
        qp[2 * dd] = pack_f32x2(v.x, v.y);
        qp[2 * dd + 1] = pack_f32x2(v.z, v.w);
      }
      unsigned long long Up[8][2][2];
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int u = 0; u < 2; ++u) Up[m][u][0] = Up[m][u][1] = 0ull;   // two +0.0f
      const int sw = g & 1;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const float* row = s_w1 + (g * 8 + d) * 256;
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 w4 = ld4(row + 4 * ((8 * m + 2 * t + u) ^ sw));
            Up[m][u][0] = fma_f32x2(pack_f32x2(w4.x, w4.y), qp[d], Up[m][u][0]);
            Up[m][u][1] = fma_f32x2(pack_f32x2(w4.z, w4.w), qp[d], Up[m][u][1]);
          }
      }
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          unpack_f32x2(Up[m][u][0], Uf[m][u].x, Uf[m][u].y);
          unpack_f32x2(Up[m][u][1], Uf[m][u].z, Uf[m][u].w);
        }
    }
    const float* pik = pist + 4 * t;
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[nt][k] = 0.f;

#pragma unroll
    for (int pass = 0; pass < 4 / RP; ++pass) {
      float4 a[RP][8];
      float rstd[RP];
#pragma unroll
      for (int rr = 0; rr < RP; ++rr) {
        const int nt = RP * pass + rr;
        const int e = g + 8 * nt;
        const float* pj = p.pj_k + (size_t)M.j[e] * CBG_H + 4 * t;
#pragma unroll
        for (int m = 0; m < 8; ++m) a[rr][m] = add4(ld4(pik + 16 * m), ldg4(pj + 16 * m));
        if (8 * nt + 7 < nst) {                                   // warp-uniform: the whole block is static
          const float* r = rc + M.slot[e] * CBG_H + 4 * t;
#pragma unroll
          for (int m = 0; m < 8; ++m) a[rr][m] = add4(a[rr][m], ldg4(r + 16 * m));
        } else {
          const int tt = M.t[e];
          const float* cc = s_c + tt * CBG_H + 4 * t;
#pragma unroll
          for (int m = 0; m < 8; ++m) a[rr][m] = add4(a[rr][m], ld4(cc + 16 * m));
          const float* w = s_wrf + tt * (CBG_NRBF * CBG_H) + 4 * t;
#pragma unroll 2
          for (int mm = 0; mm < CBG_NRBF; ++mm) {
            const float gv = M.g[mm][e];
#pragma unroll
            for (int m = 0; m < 8; ++m) fma4(a[rr][m], ld4(w + mm * CBG_H + 16 * m), gv);
          }
        }
        // LayerNorm statistics: the row lives in the 4 lanes of the quad
        float s = 0.f;
#pragma unroll
        for (int m = 0; m < 8; ++m) s += (a[rr][m].x + a[rr][m].y) + (a[rr][m].z + a[rr][m].w);
        s += __shfl_xor_sync(CBG_FULL, s, 1);
        s += __shfl_xor_sync(CBG_FULL, s, 2);
        const float mean = s * (1.f / 128.f);
        float v = 0.f;
#pragma unroll
        for (int m = 0; m < 8; ++m) { a[rr][m] = add4s(a[rr][m], -mean); v += dot4(a[rr][m], a[rr][m]); }
        v += __shfl_xor_sync(CBG_FULL, v, 1);
        v += __shfl_xor_sync(CBG_FULL, v, 2);
        rstd[rr] = 1.f / sqrtf(v * (1.f / 128.f) + 1e-5f);
      }
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const float4 gamma = ld4(s_ln + 16 * m + 4 * t), beta = ld4(s_ln + 128 + 16 * m + 4 * t);
#pragma unroll
        for (int rr = 0; rr < RP; ++rr) a[rr][m] = ln_relu4(a[rr][m], rstd[rr], gamma, beta);
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          // k-tile 2m+u: slot t <-> feature 16m+4t+2u, slot t+4 <-> feature 16m+4t+2u+1
          const float af[4] = {Uf[m][u].x, Uf[m][u].y, Uf[m][u].z, Uf[m][u].w};
          unsigned ah[4], al[4];
          split_frag(af, ah, al);
#pragma unroll
          for (int rr = 0; rr < RP; ++rr) {
            const float bf[2] = {comp4(a[rr][m], 2 * u), comp4(a[rr][m], 2 * u + 1)};
            unsigned bh[2], bl[2];
            split_frag(bf, bh, bl);
            mma3(acc[RP * pass + rr], ah, al, bh, bl);
          }
        }
      }
    }
    // acc[nt][0..1]: head g, edges 8nt + 2t, 8nt + 2t + 1; acc[nt][2..3]: head g + 8.  Softmax over the 32 edges
    // of a head = 8 values in this lane x the 4 lanes of the quad.
    float* wout = p.w + (size_t)i * (CBG_KMAX * CBG_HEADS);
#pragma unroll
    for (int hs = 0; hs < 2; ++hs) {
      float l[8];
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const bool valid = (vmask >> (8 * nt + 2 * t + v)) & 1u;
          l[2 * nt + v] = valid ? acc[nt][2 * hs + v] : -INFINITY;
          mx = fmaxf(mx, l[2 * nt + v]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { l[k] = (mx == -INFINITY) ? 0.f : expf(l[k] - mx); sum += l[k]; }
      sum += __shfl_xor_sync(CBG_FULL, sum, 1);
      sum += __shfl_xor_sync(CBG_FULL, sum, 2);
      const float inv = (sum > 0.f) ? sum : 1.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int e = 8 * nt + 2 * t + v;
          wout[e * CBG_HEADS + g + 8 * hs] = (l[2 * nt + v] / inv) * M.ew[e];
        }
    }
    __syncwarp();   // M is rewritten by the next node's setup
  }
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) x2h_v_mma_kernel(EdgeArgs p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int kHead = 4 * 20 * 128 + 4 * 128 + 256;     // WRF | C | LN
  const float* s_wrf = smem;
  const float* s_c = s_wrf + 4 * 20 * 128;
  const float* s_ln = s_c + 4 * 128;
  const float* s_w1 = s_ln + 256;                          // fragment layout, see block_copy_w1*_frag
  const float* s_b1 = s_w1 + 128 * 128;
  const float* s_rbf = s_b1 + 128;
  EdgeMeta* metas = reinterpret_cast<EdgeMeta*>(smem + kX2hVFloats);
  {
    const float* src = p.layer + kOffX2hV;
    block_copy_f4(smem, src, kHead);
    block_copy_w1v_frag(smem + kHead, src + kHead);
    block_copy_f4(smem + kHead + 128 * 128, src + kHead + 128 * 128, 128 + 32);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  EdgeMeta& M = metas[warp];
  float* pist = reinterpret_cast<float*>(metas + kWarps) + warp * 128;   // per-warp staging of Pi_i

  const int n_list = list_length(p);
  NodeSched sch;
  for (int n = sch.first(p, blockIdx.x * kWarps + warp, gridDim.x * kWarps, lane); n < n_list; n = sch.nn) {
    sch.draw(n, lane);
    const int i = p.node_idx ? p.node_idx[n] : n;
    edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf, p.fstat);
    const float* rc = p.rc_v ? p.rc_v + (size_t)i * (CBG_KMAX * CBG_H) : nullptr;
    const int nst = rc ? __popc(__ballot_sync(CBG_FULL, M.slot[lane] >= 0)) : 0;
    {
      const int nn = sch.next(lane);
      if (nn < n_list) prefetch_rc(p.rc_v, p.node_idx ? p.node_idx[nn] : nn, lane);
    }
    // attention weights (alpha * e_w) as A fragments: row = head (g, g+8), column = edge (8kt+t, 8kt+t+4)
    float wf[4][4];
    float sw0 = 0.f, sw1 = 0.f;
    {
      const float* wsrc = p.w + (size_t)i * (CBG_KMAX * CBG_HEADS);
#pragma unroll
      for (int kt = 0; kt < 4; ++kt) {
        const float* r0 = wsrc + (8 * kt + t) * CBG_HEADS + g;
        wf[kt][0] = r0[0];
        wf[kt][1] = r0[8];
        wf[kt][2] = r0[4 * CBG_HEADS];
        wf[kt][3] = r0[4 * CBG_HEADS + 8];
        sw0 += wf[kt][0] + wf[kt][2];
        sw1 += wf[kt][1] + wf[kt][3];
      }
      sw0 += __shfl_xor_sync(CBG_FULL, sw0, 1);
      sw0 += __shfl_xor_sync(CBG_FULL, sw0, 2);
      sw1 += __shfl_xor_sync(CBG_FULL, sw1, 1);
      sw1 += __shfl_xor_sync(CBG_FULL, sw1, 2);
    }
    st4(pist + 4 * lane, ldg4(p.pi_v + (size_t)i * CBG_H + 4 * lane));     // Pi row staged in shared memory, re-read per k-tile
    __syncwarp();
    // S[head][feature]: acc[c][q][0..1] = head g, features 32c + 8t + q and 32c + 8t + 4 + q; [2..3] = head g + 8
    float acc[4][4][4];
#pragma unroll
    for (int c = 0; c < 4; ++c)
#pragma unroll
      for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[c][q][k] = 0.f;

#pragma unroll
    for (int kt = 0; kt < 4; ++kt) {
      // the two edges of this lane (k slots t and t+4 of the k-tile) are kept as packed pairs per feature:
      // a2[c][q] = (edge e0, edge e1) of feature 32c + 4g + q  ==  the B fragment (b0, b1) of n-tile (c, q)
      const int e0 = 8 * kt + t, e1 = e0 + 4;
      float2 a2[4][4];
      {
        const float* pj0 = p.pj_v + (size_t)M.j[e0] * CBG_H + 4 * g;
        const float* pj1 = p.pj_v + (size_t)M.j[e1] * CBG_H + 4 * g;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 x0 = ldg4(pj0 + 32 * c), x1 = ldg4(pj1 + 32 * c), pi4 = ld4(pist + 4 * g + 32 * c);
          a2[c][0] = make_float2(pi4.x + x0.x, pi4.x + x1.x);
          a2[c][1] = make_float2(pi4.y + x0.y, pi4.y + x1.y);
          a2[c][2] = make_float2(pi4.z + x0.z, pi4.z + x1.z);
          a2[c][3] = make_float2(pi4.w + x0.w, pi4.w + x1.w);
        }
      }
      if (8 * kt + 7 < nst) {                                   // warp-uniform: the whole block is static
        const float* r0 = rc + M.slot[e0] * CBG_H + 4 * g;
        const float* r1 = rc + M.slot[e1] * CBG_H + 4 * g;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 x0 = ldg4(r0 + 32 * c), x1 = ldg4(r1 + 32 * c);
          a2[c][0].x += x0.x; a2[c][0].y += x1.x;
          a2[c][1].x += x0.y; a2[c][1].y += x1.y;
          a2[c][2].x += x0.z; a2[c][2].y += x1.z;
          a2[c][3].x += x0.w; a2[c][3].y += x1.w;
        }
      } else {
        const int t0 = M.t[e0], t1 = M.t[e1];
        const float* c0p = s_c + t0 * CBG_H + 4 * g;
        const float* c1p = s_c + t1 * CBG_H + 4 * g;
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          const float4 x0 = ld4(c0p + 32 * c), x1 = ld4(c1p + 32 * c);
          a2[c][0].x += x0.x; a2[c][0].y += x1.x;
          a2[c][1].x += x0.y; a2[c][1].y += x1.y;
          a2[c][2].x += x0.z; a2[c][2].y += x1.z;
          a2[c][3].x += x0.w; a2[c][3].y += x1.w;
        }
        const float* w0 = s_wrf + t0 * (CBG_NRBF * CBG_H) + 4 * g;
        if (t0 == t1) {                                          // same weight rows for both edges: packed FMAs
#pragma unroll 4
          for (int mm = 0; mm < CBG_NRBF; ++mm) {
            const float2 gp = make_float2(M.g[mm][e0], M.g[mm][e1]);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 wv = ld4(w0 + mm * CBG_H + 32 * c);
              a2[c][0] = __ffma2_rn(gp, make_float2(wv.x, wv.x), a2[c][0]);
              a2[c][1] = __ffma2_rn(gp, make_float2(wv.y, wv.y), a2[c][1]);
              a2[c][2] = __ffma2_rn(gp, make_float2(wv.z, wv.z), a2[c][2]);
              a2[c][3] = __ffma2_rn(gp, make_float2(wv.w, wv.w), a2[c][3]);
            }
          }
        } else {
          const float* w1 = s_wrf + t1 * (CBG_NRBF * CBG_H) + 4 * g;
#pragma unroll 1
          for (int mm = 0; mm < CBG_NRBF; ++mm) {
            const float g0 = M.g[mm][e0], g1 = M.g[mm][e1];
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const float4 wa = ld4(w0 + mm * CBG_H + 32 * c), wb = ld4(w1 + mm * CBG_H + 32 * c);
              a2[c][0].x = fmaf(wa.x, g0, a2[c][0].x); a2[c][0].y = fmaf(wb.x, g1, a2[c][0].y);
              a2[c][1].x = fmaf(wa.y, g0, a2[c][1].x); a2[c][1].y = fmaf(wb.y, g1, a2[c][1].y);
              a2[c][2].x = fmaf(wa.z, g0, a2[c][2].x); a2[c][2].y = fmaf(wb.z, g1, a2[c][2].y);
              a2[c][3].x = fmaf(wa.w, g0, a2[c][3].x); a2[c][3].y = fmaf(wb.w, g1, a2[c][3].y);
            }
          }
        }
      }
      // LayerNorm statistics of both edges at once; a row lives in the 8 lanes with the same t
      float2 s2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c)
        s2 = __fadd2_rn(s2, __fadd2_rn(__fadd2_rn(a2[c][0], a2[c][1]), __fadd2_rn(a2[c][2], a2[c][3])));
#pragma unroll
      for (int sh = 4; sh <= 16; sh <<= 1) {
        s2.x += __shfl_xor_sync(CBG_FULL, s2.x, sh);
        s2.y += __shfl_xor_sync(CBG_FULL, s2.y, sh);
      }
      const float2 nmean = make_float2(-s2.x * (1.f / 128.f), -s2.y * (1.f / 128.f));
      float2 v2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          a2[c][q] = __fadd2_rn(a2[c][q], nmean);
          v2 = __ffma2_rn(a2[c][q], a2[c][q], v2);
        }
#pragma unroll
      for (int sh = 4; sh <= 16; sh <<= 1) {
        v2.x += __shfl_xor_sync(CBG_FULL, v2.x, sh);
        v2.y += __shfl_xor_sync(CBG_FULL, v2.y, sh);
      }
      const float2 rstd2 = make_float2(1.f / sqrtf(v2.x * (1.f / 128.f) + 1e-5f), 1.f / sqrtf(v2.y * (1.f / 128.f) + 1e-5f));
      unsigned ah[4], al[4];
      split_frag(wf[kt], ah, al);
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float4 gamma = ld4(s_ln + 32 * c + 4 * g), beta = ld4(s_ln + 128 + 32 * c + 4 * g);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          // n-tile (c, q): column g <-> feature 32c + 4g + q ; k slots t, t+4 <-> edges e0, e1
          const float gq = comp4(gamma, q), bq = comp4(beta, q);
          const float2 y = __ffma2_rn(__fmul2_rn(a2[c][q], rstd2), make_float2(gq, gq), make_float2(bq, bq));
          const float bf[2] = {fmaxf(y.x, 0.f), fmaxf(y.y, 0.f)};
          unsigned bh[2], bl[2];
          split_frag(bf, bh, bl);
          mma3(acc[c][q], ah, al, bh, bl);
        }
      }
    }
    // out[f'] = W1v[f'][:] . S[head(f')][:] : this lane owns heads g, g+8 and the features 32c + 8t + 4v + q,
    // so it forms 16 partial sums (2 heads x 8 rows of W1v), which are then reduced over the quad.
    float part[16];
    const int sw = g & 1;
#pragma unroll
    for (int sel = 0; sel < 2; ++sel)
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const float* wrow = s_w1 + ((g + 8 * sel) * 8 + d) * CBG_H;
        float2 t2 = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int hq = 0; hq < 2; ++hq) {
            // chunk = {(q=2hq, v=0), (q=2hq, v=1), (q=2hq+1, v=0), (q=2hq+1, v=1)}; accumulator pairs are (v=0, v=1)
            const float4 w4 = ld4(wrow + 4 * ((8 * c + 2 * t + hq) ^ sw));
            t2 = __ffma2_rn(make_float2(w4.x, w4.y), make_float2(acc[c][2 * hq][2 * sel], acc[c][2 * hq][2 * sel + 1]), t2);
            t2 = __ffma2_rn(make_float2(w4.z, w4.w), make_float2(acc[c][2 * hq + 1][2 * sel], acc[c][2 * hq + 1][2 * sel + 1]), t2);
          }
        part[sel * 8 + d] = t2.x + t2.y;
      }
    {   // quad bit 1 <-> head select
      const bool up = (t & 2) != 0;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float send = up ? part[k] : part[k + 8];
        const float keep = up ? part[k + 8] : part[k];
        part[k] = keep + __shfl_xor_sync(CBG_FULL, send, 2);
      }
    }
    {   // quad bit 0 <-> upper / lower 4 rows of the head
      const bool up = (t & 1) != 0;
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float send = up ? part[k] : part[k + 4];
        const float keep = up ? part[k + 4] : part[k];
        part[k] = keep + __shfl_xor_sync(CBG_FULL, send, 1);
      }
    }
    {
      const int sel = (t >> 1) & 1;
      const int f0 = (g + 8 * sel) * 8 + 4 * (t & 1);
      const float swh = sel ? sw1 : sw0;
      float* hrow = p.h + (size_t)i * CBG_H + f0;
      const float4 hin = ld4(hrow), b1 = ld4(s_b1 + f0);
      st4(hrow, make_float4(hin.x + (part[0] + b1.x * swh), hin.y + (part[1] + b1.y * swh),
                            hin.z + (part[2] + b1.z * swh), hin.w + (part[3] + b1.w * swh)));
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// x2h_k, second tensor-core variant: the RBF mat-vec of the non-cached edges runs on the tensor cores as well.
//
// Per pass of 16 edges (positions 16p .. 16p+15; lane (g, t) owns e0 = 16p + g and e1 = e0 + 8) the first Linear is
//     pre[e][f] = Pi[f] + Pj[j_e][f] + (static e ? R[e][f] : c[type_e][f] + sum_m g_m(d_e) Wrf[type_e][m][f])
// The last term is the GEMM  G[16 edges x 24] . Wrf[type][24 x 128]  (20 RBFs padded to 3 k-tiles): one m16n8k8 tile per
// 8 features, with the rows of edges of another type (or static / padded edges) zeroed in G, once per edge type present.
// Its accumulator fragment {(e0,f), (e0,f+1), (e1,f), (e1,f+1)} IS the register layout of the activations (act[m][h][4],
// f = 16m + 4t + 2h), so the MMA accumulates straight into them; the same registers are then the B fragments of the
// head contraction.  Static edges are decided per lane (no more "whole block dynamic because of one edge").
// Wrf lives in shared memory as B fragments: img[type][kt][n-tile (m,h)][lane][2] (see block_copy_wrf_frag).
constexpr int kWimgFloats = 4 * 3 * 16 * 32 * 2;                                  // 12288 floats = 48 KB
constexpr int kX2hK2Floats = kWimgFloats + 4 * 128 + 256 + 128 * 128 + 32;        // WIMG | C | LN | W1 (fragments) | RBF
constexpr int x2hk_mma2_smem(int w) { return kX2hK2Floats * 4 + w * ((int)sizeof(EdgeMeta) + 256 * 4); }

// B fragments of the RBF GEMM: b_j (j = 0, 1) of lane (g', t') for k-tile kt and n-tile (m, h) is
// Wrf[type][8kt + t' + 4j][16m + 4(g' >> 1) + 2h + (g' & 1)]  (0 for the padded RBF rows 20..23)
__device__ __forceinline__ void block_copy_wrf_frag(float* dst, const float* __restrict__ wrf /*[4][20][128]*/) {
  for (int idx = threadIdx.x; idx < kWimgFloats; idx += blockDim.x) {
    const int j = idx & 1, ln = (idx >> 1) & 31, nt = (idx >> 6) & 15, kt = (idx >> 10) % 3, ty = idx / (3 << 10);
    const int gp = ln >> 2, tp = ln & 3, m = nt >> 1, h = nt & 1;
    const int rbf = 8 * kt + tp + 4 * j, f = 16 * m + 4 * (gp >> 1) + 2 * h + (gp & 1);
    dst[idx] = rbf < CBG_NRBF ? __ldg(wrf + (ty * CBG_NRBF + rbf) * CBG_H + f) : 0.f;
  }
}

template <int kWarps>
__global__ void __launch_bounds__(kWarps * 32, 1) x2h_k_mma2_kernel(EdgeArgs p) {
  extern __shared__ __align__(16) float smem[];
  const float* s_wimg = smem;
  const float* s_c = s_wimg + kWimgFloats;
  const float* s_ln = s_c + 4 * 128;
  const float* s_w1 = s_ln + 256;                          // fragment layout, see block_copy_w1k_frag
  const float* s_rbf = s_w1 + 128 * 128;
  EdgeMeta* metas = reinterpret_cast<EdgeMeta*>(smem + kX2hK2Floats);
  {
    const float* src = p.layer + kOffX2hK;                 // blob: WRF | C | LN | W1 | RBF
    block_copy_wrf_frag(smem, src);
    block_copy_f4(smem + kWimgFloats, src + 4 * 20 * 128, 4 * 128 + 256);
    block_copy_w1k_frag(smem + kWimgFloats + 4 * 128 + 256, src + 4 * 20 * 128 + 4 * 128 + 256);
    block_copy_f4(smem + kWimgFloats + 4 * 128 + 256 + 128 * 128, src + 4 * 20 * 128 + 4 * 128 + 256 + 128 * 128, 32);
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  EdgeMeta& M = metas[warp];
  float* qst = reinterpret_cast<float*>(metas + kWarps) + warp * 256;   // per-warp staging of q_i (fragment order) | Pi_i
  float* pist = qst + 128;

  const int n_list = list_length(p);
  NodeSched sch;
  for (int n = sch.first(p, blockIdx.x * kWarps + warp, gridDim.x * kWarps, lane); n < n_list; n = sch.nn) {
    sch.draw(n, lane);
    const int i = p.node_idx ? p.node_idx[n] : n;
    const unsigned vmask = edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf, p.fstat);
    const float* rc = p.rc_k ? p.rc_k + (size_t)i * (CBG_KMAX * CBG_H) : nullptr;
    const int nst = rc ? __popc(__ballot_sync(CBG_FULL, M.slot[lane] >= 0)) : 0;   // static edges occupy positions 0 .. nst-1
    {
      const int nn = sch.next(lane);
      if (nn < n_list) prefetch_rc(p.rc_k, p.node_idx ? p.node_idx[nn] : nn, lane);
    }
    // query-folded key matrix as A fragments (same as x2h_k_mma_kernel)
    float4 Uf[8][2];
    {
      {
        const float4 q4 = ldg4(p.q + (size_t)i * CBG_H + 4 * lane);
        const int hd = lane >> 1, d0 = 4 * (lane & 1);
        float* dst = qst + ((hd & 7) * 8 + d0) * 2 + (hd >> 3);
        dst[0] = q4.x; dst[2] = q4.y; dst[4] = q4.z; dst[6] = q4.w;
        st4(pist + 4 * lane, ldg4(p.pi_k + (size_t)i * CBG_H + 4 * lane));
      }
      __syncwarp();
      unsigned long long qp[8];
#pragma unroll
      for (int dd = 0; dd < 4; ++dd) {
        const float4 v = ld4(qst + g * 16 + 4 * dd);
        qp[2 * dd] = pack_f32x2(v.x, v.y);
        qp[2 * dd + 1] = pack_f32x2(v.z, v.w);
      }
      unsigned long long Up[8][2][2];
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int u = 0; u < 2; ++u) Up[m][u][0] = Up[m][u][1] = 0ull;
      const int sw = g & 1;
#pragma unroll
      for (int d = 0; d < 8; ++d) {
        const float* row = s_w1 + (g * 8 + d) * 256;
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            const float4 w4 = ld4(row + 4 * ((8 * m + 2 * t + u) ^ sw));
            Up[m][u][0] = fma_f32x2(pack_f32x2(w4.x, w4.y), qp[d], Up[m][u][0]);
            Up[m][u][1] = fma_f32x2(pack_f32x2(w4.z, w4.w), qp[d], Up[m][u][1]);
          }
      }
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int u = 0; u < 2; ++u) {
          unpack_f32x2(Up[m][u][0], Uf[m][u].x, Uf[m][u].y);
          unpack_f32x2(Up[m][u][1], Uf[m][u].z, Uf[m][u].w);
        }
    }
    float acc[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int k = 0; k < 4; ++k) acc[nt][k] = 0.f;

#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
      const int e0 = 16 * pass + g, e1 = e0 + 8;
      const bool st0 = e0 < nst, st1 = e1 < nst;
      // act[m][h] = {(e0,f), (e0,f+1), (e1,f), (e1,f+1)}, f = 16m + 4t + 2h
      float act[8][2][4];
      {
        const float* pj0 = p.pj_k + (size_t)M.j[e0] * CBG_H + 4 * t;
        const float* pj1 = p.pj_k + (size_t)M.j[e1] * CBG_H + 4 * t;
        // third term: the R-cache row of a static edge, the type constant of any other edge
        const float* r0 = st0 ? rc + M.slot[e0] * CBG_H + 4 * t : nullptr;
        const float* r1 = st1 ? rc + M.slot[e1] * CBG_H + 4 * t : nullptr;
        const float* c0 = s_c + (st0 ? 0 : M.t[e0]) * CBG_H + 4 * t;     // (M.t is not written by the static fast path)
        const float* c1 = s_c + (st1 ? 0 : M.t[e1]) * CBG_H + 4 * t;
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          const float4 pi4 = ld4(pist + 4 * t + 16 * m);
          const float4 x0 = ldg4(pj0 + 16 * m), x1 = ldg4(pj1 + 16 * m);
          const float4 y0 = st0 ? ldg4(r0 + 16 * m) : ld4(c0 + 16 * m);
          const float4 y1 = st1 ? ldg4(r1 + 16 * m) : ld4(c1 + 16 * m);
          act[m][0][0] = (pi4.x + x0.x) + y0.x; act[m][0][1] = (pi4.y + x0.y) + y0.y;
          act[m][0][2] = (pi4.x + x1.x) + y1.x; act[m][0][3] = (pi4.y + x1.y) + y1.y;
          act[m][1][0] = (pi4.z + x0.z) + y0.z; act[m][1][1] = (pi4.w + x0.w) + y0.w;
          act[m][1][2] = (pi4.z + x1.z) + y1.z; act[m][1][3] = (pi4.w + x1.w) + y1.w;
        }
      }
      if (__any_sync(CBG_FULL, !st0 || !st1)) {          // warp-uniform: some edge of this pass needs the RBF term
        const int t0 = st0 ? -1 : M.t[e0], t1 = st1 ? -1 : M.t[e1];
#pragma unroll 1
        for (int ty = 0; ty < CBG_NTYPE; ++ty) {
          if (!__any_sync(CBG_FULL, t0 == ty || t1 == ty)) continue;
          const float m0 = t0 == ty ? 1.f : 0.f, m1 = t1 == ty ? 1.f : 0.f;
          unsigned gh[3][4], gl[3][4];
#pragma unroll
          for (int kt = 0; kt < 3; ++kt) {
            const int ra = 8 * kt + t, rb = ra + 4;        // rb >= 20 only for kt == 2: padded rows
            const float gf[4] = {M.g[ra][e0] * m0, M.g[ra][e1] * m1,
                                 kt < 2 ? M.g[kt < 2 ? rb : 0][e0] * m0 : 0.f, kt < 2 ? M.g[kt < 2 ? rb : 0][e1] * m1 : 0.f};
            split_frag(gf, gh[kt], gl[kt]);
          }
          const float* img = s_wimg + ty * (3 * 16 * 64) + 2 * lane;
#pragma unroll
          for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
              for (int kt = 0; kt < 3; ++kt) {
                const float2 wv = *reinterpret_cast<const float2*>(img + (kt * 16 + 2 * m + h) * 64);
                const float bf[2] = {wv.x, wv.y};
                unsigned bh[2], bl[2];
                split_frag(bf, bh, bl);
                mma3(act[m][h], gh[kt], gl[kt], bh, bl);
              }
        }
      }
      // LayerNorm statistics of both edges; a row lives in the 4 lanes of the quad
      float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          s01 = __fadd2_rn(s01, make_float2(act[m][h][0], act[m][h][1]));
          s23 = __fadd2_rn(s23, make_float2(act[m][h][2], act[m][h][3]));
        }
      float sa = s01.x + s01.y, sb = s23.x + s23.y;
      sa += __shfl_xor_sync(CBG_FULL, sa, 1); sb += __shfl_xor_sync(CBG_FULL, sb, 1);
      sa += __shfl_xor_sync(CBG_FULL, sa, 2); sb += __shfl_xor_sync(CBG_FULL, sb, 2);
      const float na = -sa * (1.f / 128.f), nb = -sb * (1.f / 128.f);
      float2 v01 = make_float2(0.f, 0.f), v23 = make_float2(0.f, 0.f);
#pragma unroll
      for (int m = 0; m < 8; ++m)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 d01 = __fadd2_rn(make_float2(act[m][h][0], act[m][h][1]), make_float2(na, na));
          const float2 d23 = __fadd2_rn(make_float2(act[m][h][2], act[m][h][3]), make_float2(nb, nb));
          act[m][h][0] = d01.x; act[m][h][1] = d01.y; act[m][h][2] = d23.x; act[m][h][3] = d23.y;
          v01 = __ffma2_rn(d01, d01, v01);
          v23 = __ffma2_rn(d23, d23, v23);
        }
      float va = v01.x + v01.y, vb = v23.x + v23.y;
      va += __shfl_xor_sync(CBG_FULL, va, 1); vb += __shfl_xor_sync(CBG_FULL, vb, 1);
      va += __shfl_xor_sync(CBG_FULL, va, 2); vb += __shfl_xor_sync(CBG_FULL, vb, 2);
      const float ra_ = 1.f / sqrtf(va * (1.f / 128.f) + 1e-5f), rb_ = 1.f / sqrtf(vb * (1.f / 128.f) + 1e-5f);
#pragma unroll
      for (int m = 0; m < 8; ++m) {
        const float4 gamma = ld4(s_ln + 16 * m + 4 * t), beta = ld4(s_ln + 128 + 16 * m + 4 * t);
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const float2 gm = h ? make_float2(gamma.z, gamma.w) : make_float2(gamma.x, gamma.y);
          const float2 bt = h ? make_float2(beta.z, beta.w) : make_float2(beta.x, beta.y);
          const float2 y0 = __ffma2_rn(__fmul2_rn(make_float2(act[m][h][0], act[m][h][1]), make_float2(ra_, ra_)), gm, bt);
          const float2 y1 = __ffma2_rn(__fmul2_rn(make_float2(act[m][h][2], act[m][h][3]), make_float2(rb_, rb_)), gm, bt);
          const float af[4] = {h ? Uf[m][1].x : Uf[m][0].x, h ? Uf[m][1].y : Uf[m][0].y, h ? Uf[m][1].z : Uf[m][0].z, h ? Uf[m][1].w : Uf[m][0].w};
          unsigned ah[4], al[4];
          split_frag(af, ah, al);
          const float b0[2] = {fmaxf(y0.x, 0.f), fmaxf(y0.y, 0.f)}, b1[2] = {fmaxf(y1.x, 0.f), fmaxf(y1.y, 0.f)};
          unsigned bh[2], bl[2];
          split_frag(b0, bh, bl);
          mma3(acc[2 * pass], ah, al, bh, bl);
          split_frag(b1, bh, bl);
          mma3(acc[2 * pass + 1], ah, al, bh, bl);
        }
      }
    }
    // softmax + store: identical to x2h_k_mma_kernel
    float* wout = p.w + (size_t)i * (CBG_KMAX * CBG_HEADS);
#pragma unroll
    for (int hs = 0; hs < 2; ++hs) {
      float l[8];
      float mx = -INFINITY;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const bool valid = (vmask >> (8 * nt + 2 * t + v)) & 1u;
          l[2 * nt + v] = valid ? acc[nt][2 * hs + v] : -INFINITY;
          mx = fmaxf(mx, l[2 * nt + v]);
        }
      mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 2));
      float sum = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) { l[k] = (mx == -INFINITY) ? 0.f : expf(l[k] - mx); sum += l[k]; }
      sum += __shfl_xor_sync(CBG_FULL, sum, 1);
      sum += __shfl_xor_sync(CBG_FULL, sum, 2);
      const float inv = (sum > 0.f) ? sum : 1.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int e = 8 * nt + 2 * t + v;
          wout[e * CBG_HEADS + g + 8 * hs] = (l[2 * nt + v] / inv) * M.ew[e];
        }
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------------------
// H2X on the tensor cores, two warps per generated node.
//
// Every in-edge of a generated node is "dynamic" (its end point moves), so both edge MLPs of H2X take the RBF path for
// all 32 edges; with one warp per node the launch is one long latency chain (1 536 nodes on 1 776 warp slots at c2).
// Here a node is handled by a PAIR of warps of the same CTA:
//   K warp: edge_setup, query folding, k-MLP (two 16-edge passes exactly like x2h_k_mma2_kernel: RBF mat-vec and head
//           contraction as MMAs), softmax -> alpha * e_w into the pair's shared buffer
//   V warp: v-MLP with the same pass code (A fragments = xv_func's second Linear [16 heads x 128], from shared memory),
//           then dx_i = 1/16 sum_{e,hd} alpha e_w (W1xv a + b1)[e][hd] (x_i - x_j)
// The two meet at named barriers (one id per pair): M ready -> alpha ready -> node done.
constexpr int kH2xPairK = kWimgFloats + 4 * 128 + 256 + 128 * 128;          // K: WIMG | C | LN | W1 (U fragments)
constexpr int kH2xPairV = kWimgFloats + 4 * 128 + 256 + 16 * 128 + 32;      // V: WIMG | C | LN | W1 (A fragments) | B1
constexpr int kH2xPairFloats = kH2xPairK + kH2xPairV + 32;                  // + RBF
constexpr int kPairScratchFloats = 128 + 128 + 128 + 32 * 16 + 4;           // q (fragment order) | Pi_k | Pi_v | alpha * e_w | next node
constexpr int h2x_pair_smem(int pairs) { return kH2xPairFloats * 4 + pairs * ((int)sizeof(EdgeMeta) + kPairScratchFloats * 4); }

// A fragments of xv_func's second Linear: frag[(m*2+h)*32 + lane] = {W[g][f], W[g+8][f], W[g][f+1], W[g+8][f+1]}, f = 16m+4t+2h
__device__ __forceinline__ void block_copy_w1xv_frag(float* dst, const float* __restrict__ w /*[16][128]*/) {
  for (int idx = threadIdx.x; idx < 16 * 32; idx += blockDim.x) {
    const int ln = idx & 31, mh = idx >> 5, gp = ln >> 2, tp = ln & 3;
    const int f = 16 * (mh >> 1) + 4 * tp + 2 * (mh & 1);
    st4(dst + 4 * idx, make_float4(__ldg(w + gp * CBG_H + f), __ldg(w + (gp + 8) * CBG_H + f),
                                   __ldg(w + gp * CBG_H + f + 1), __ldg(w + (gp + 8) * CBG_H + f + 1)));
  }
}

__device__ __forceinline__ void pair_barrier(int id) { asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory"); }

struct RegFrag {          // A fragments held in registers (query-folded key matrix)
  const float4 (&U)[8][2];
  __device__ __forceinline__ float4 get(int m, int h) const { return h ? U[m][1] : U[m][0]; }
};
struct SmemFrag {         // A fragments in shared memory, one LDS.128 per k-tile
  const float* img;
  int lane;
  __device__ __forceinline__ float4 get(int m, int h) const { return ld4(img + ((m * 2 + h) * 32 + lane) * 4); }
};

// One 16-edge pass of an edge MLP whose edges are all non-cached (see x2h_k_mma2_kernel for the layout):
// acc0 / acc1 += A . relu(LN(Pi + Pj[j] + c[type] + Wrf[type] g))^T for the edge blocks 2*pass and 2*pass + 1.
template <class AFrag>
__device__ __forceinline__ void mlp_pass16_dynamic(const EdgeMeta& M, int pass, int g, int t, int lane,
                                                   const float* pist, const float* __restrict__ pj_plane,
                                                   const float* s_c, const float* s_wimg, const float* s_ln,
                                                   const AFrag& A, float (&acc0)[4], float (&acc1)[4]) {
  const int e0 = 16 * pass + g, e1 = e0 + 8;
  const int t0 = M.t[e0], t1 = M.t[e1];
  float act[8][2][4];
  {
    const float* pj0 = pj_plane + (size_t)M.j[e0] * CBG_H + 4 * t;
    const float* pj1 = pj_plane + (size_t)M.j[e1] * CBG_H + 4 * t;
    const float* c0 = s_c + t0 * CBG_H + 4 * t;
    const float* c1 = s_c + t1 * CBG_H + 4 * t;
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float4 pi4 = ld4(pist + 4 * t + 16 * m);
      const float4 x0 = ldg4(pj0 + 16 * m), x1 = ldg4(pj1 + 16 * m);
      const float4 y0 = ld4(c0 + 16 * m), y1 = ld4(c1 + 16 * m);
      act[m][0][0] = (pi4.x + x0.x) + y0.x; act[m][0][1] = (pi4.y + x0.y) + y0.y;
      act[m][0][2] = (pi4.x + x1.x) + y1.x; act[m][0][3] = (pi4.y + x1.y) + y1.y;
      act[m][1][0] = (pi4.z + x0.z) + y0.z; act[m][1][1] = (pi4.w + x0.w) + y0.w;
      act[m][1][2] = (pi4.z + x1.z) + y1.z; act[m][1][3] = (pi4.w + x1.w) + y1.w;
    }
  }
#pragma unroll 1
  for (int ty = 0; ty < CBG_NTYPE; ++ty) {
    if (!__any_sync(CBG_FULL, t0 == ty || t1 == ty)) continue;
    const float m0 = t0 == ty ? 1.f : 0.f, m1 = t1 == ty ? 1.f : 0.f;
    unsigned gh[3][4], gl[3][4];
#pragma unroll
    for (int kt = 0; kt < 3; ++kt) {
      const int ra = 8 * kt + t;
      const float gf[4] = {M.g[ra][e0] * m0, M.g[ra][e1] * m1,
                           kt < 2 ? M.g[kt < 2 ? ra + 4 : 0][e0] * m0 : 0.f, kt < 2 ? M.g[kt < 2 ? ra + 4 : 0][e1] * m1 : 0.f};
      split_frag(gf, gh[kt], gl[kt]);
    }
    const float* img = s_wimg + ty * (3 * 16 * 64) + 2 * lane;
#pragma unroll
    for (int m = 0; m < 8; ++m)
#pragma unroll
      for (int h = 0; h < 2; ++h)
#pragma unroll
        for (int kt = 0; kt < 3; ++kt) {
          const float2 wv = *reinterpret_cast<const float2*>(img + (kt * 16 + 2 * m + h) * 64);
          const float bf[2] = {wv.x, wv.y};
          unsigned bh[2], bl[2];
          split_frag(bf, bh, bl);
          mma3(act[m][h], gh[kt], gl[kt], bh, bl);
        }
  }
  float2 s01 = make_float2(0.f, 0.f), s23 = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < 8; ++m)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      s01 = __fadd2_rn(s01, make_float2(act[m][h][0], act[m][h][1]));
      s23 = __fadd2_rn(s23, make_float2(act[m][h][2], act[m][h][3]));
    }
  float sa = s01.x + s01.y, sb = s23.x + s23.y;
  sa += __shfl_xor_sync(CBG_FULL, sa, 1); sb += __shfl_xor_sync(CBG_FULL, sb, 1);
  sa += __shfl_xor_sync(CBG_FULL, sa, 2); sb += __shfl_xor_sync(CBG_FULL, sb, 2);
  const float na = -sa * (1.f / 128.f), nb = -sb * (1.f / 128.f);
  float2 v01 = make_float2(0.f, 0.f), v23 = make_float2(0.f, 0.f);
#pragma unroll
  for (int m = 0; m < 8; ++m)
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 d01 = __fadd2_rn(make_float2(act[m][h][0], act[m][h][1]), make_float2(na, na));
      const float2 d23 = __fadd2_rn(make_float2(act[m][h][2], act[m][h][3]), make_float2(nb, nb));
      act[m][h][0] = d01.x; act[m][h][1] = d01.y; act[m][h][2] = d23.x; act[m][h][3] = d23.y;
      v01 = __ffma2_rn(d01, d01, v01);
      v23 = __ffma2_rn(d23, d23, v23);
    }
  float va = v01.x + v01.y, vb = v23.x + v23.y;
  va += __shfl_xor_sync(CBG_FULL, va, 1); vb += __shfl_xor_sync(CBG_FULL, vb, 1);
  va += __shfl_xor_sync(CBG_FULL, va, 2); vb += __shfl_xor_sync(CBG_FULL, vb, 2);
  const float ra_ = 1.f / sqrtf(va * (1.f / 128.f) + 1e-5f), rb_ = 1.f / sqrtf(vb * (1.f / 128.f) + 1e-5f);
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float4 gamma = ld4(s_ln + 16 * m + 4 * t), beta = ld4(s_ln + 128 + 16 * m + 4 * t);
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const float2 gm = h ? make_float2(gamma.z, gamma.w) : make_float2(gamma.x, gamma.y);
      const float2 bt = h ? make_float2(beta.z, beta.w) : make_float2(beta.x, beta.y);
      const float2 y0 = __ffma2_rn(__fmul2_rn(make_float2(act[m][h][0], act[m][h][1]), make_float2(ra_, ra_)), gm, bt);
      const float2 y1 = __ffma2_rn(__fmul2_rn(make_float2(act[m][h][2], act[m][h][3]), make_float2(rb_, rb_)), gm, bt);
      const float4 a4 = A.get(m, h);
      const float af[4] = {a4.x, a4.y, a4.z, a4.w};
      unsigned ah[4], al[4];
      split_frag(af, ah, al);
      const float b0[2] = {fmaxf(y0.x, 0.f), fmaxf(y0.y, 0.f)}, b1[2] = {fmaxf(y1.x, 0.f), fmaxf(y1.y, 0.f)};
      unsigned bh[2], bl[2];
      split_frag(b0, bh, bl);
      mma3(acc0, ah, al, bh, bl);
      split_frag(b1, bh, bl);
      mma3(acc1, ah, al, bh, bl);
    }
  }
}

template <int kPairs>
__global__ void __launch_bounds__(kPairs * 64, 1) h2x_pair_kernel(EdgeArgs p) {
  extern __shared__ __align__(16) float smem[];
  float* k_wimg = smem;
  float* k_c = k_wimg + kWimgFloats;
  float* k_ln = k_c + 4 * 128;
  float* k_w1 = k_ln + 256;
  float* v_wimg = k_w1 + 128 * 128;
  float* v_c = v_wimg + kWimgFloats;
  float* v_ln = v_c + 4 * 128;
  float* v_w1 = v_ln + 256;               // A fragments of xv_func.net.3
  float* v_b1 = v_w1 + 16 * 128;
  float* s_rbf = v_b1 + 32;
  EdgeMeta* metas = reinterpret_cast<EdgeMeta*>(smem + kH2xPairFloats);
  float* scratch = reinterpret_cast<float*>(metas + kPairs);
  {
    // blob (cbg_layout.h): K_WRF | K_C | K_LN | K_W1 | V_WRF | V_C | V_LN | V_W1 [16][128] | V_B1 (32) | RBF (32)
    const float* src = p.layer + kOffH2x;
    constexpr int kWrf = 4 * 20 * 128, kCL = 4 * 128 + 256;
    block_copy_wrf_frag(k_wimg, src);
    block_copy_f4(k_c, src + kWrf, kCL);
    block_copy_w1k_frag(k_w1, src + kWrf + kCL);
    const float* vsrc = src + kWrf + kCL + 128 * 128;
    block_copy_wrf_frag(v_wimg, vsrc);
    block_copy_f4(v_c, vsrc + kWrf, kCL);
    block_copy_w1xv_frag(v_w1, vsrc + kWrf + kCL);
    block_copy_f4(v_b1, vsrc + kWrf + kCL + 16 * 128, 32 + 32);       // V_B1 | RBF
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int pair = warp >> 1, role = warp & 1;
  const int g = lane >> 2, t = lane & 3;
  EdgeMeta& M = metas[pair];
  float* qst = scratch + pair * kPairScratchFloats;
  float* pik = qst + 128;
  float* piv = pik + 128;
  float* alpha = piv + 128;                 // [32 edges][16 heads]: softmax * e_w
  const int bar = 1 + pair;
  int* next_slot = reinterpret_cast<int*>(alpha + 32 * 16);
  // node scheduling: a work counter shared by all pairs (p.ticket) or a static round-robin; the K warp draws, the pair
  // reads the result after a barrier
  int n = blockIdx.x * kPairs + pair;
  if (p.ticket != nullptr) {
    if (role == 0 && lane == 0) *next_slot = atomicAdd(p.ticket, 1);
    pair_barrier(bar);
    n = *next_slot;
  }
  while (n < p.n_nodes) {
    const int i = p.node_idx[n];
    if (role == 0) {
      // ---------------- K warp ----------------
      const unsigned vmask = edge_setup(M, i, lane, p.x4, p.nbr, p.ew, s_rbf);
      pair_barrier(bar);                                            // (1) M ready
      const int nxt = (p.ticket != nullptr && lane == 0) ? atomicAdd(p.ticket, 1) : 0;   // next node: latency hidden behind this one
      float4 Uf[8][2];
      {
        {
          const float4 q4 = ldg4(p.q + (size_t)i * CBG_H + 4 * lane);
          const int hd = lane >> 1, d0 = 4 * (lane & 1);
          float* dst = qst + ((hd & 7) * 8 + d0) * 2 + (hd >> 3);
          dst[0] = q4.x; dst[2] = q4.y; dst[4] = q4.z; dst[6] = q4.w;
          st4(pik + 4 * lane, ldg4(p.pi_k + (size_t)i * CBG_H + 4 * lane));
        }
        __syncwarp();
        unsigned long long qp[8];
#pragma unroll
        for (int dd = 0; dd < 4; ++dd) {
          const float4 v = ld4(qst + g * 16 + 4 * dd);
          qp[2 * dd] = pack_f32x2(v.x, v.y);
          qp[2 * dd + 1] = pack_f32x2(v.z, v.w);
        }
        unsigned long long Up[8][2][2];
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
          for (int u = 0; u < 2; ++u) Up[m][u][0] = Up[m][u][1] = 0ull;
        const int sw = g & 1;
#pragma unroll
        for (int d = 0; d < 8; ++d) {
          const float* row = k_w1 + (g * 8 + d) * 256;
#pragma unroll
          for (int m = 0; m < 8; ++m)
#pragma unroll
            for (int u = 0; u < 2; ++u) {
              const float4 w4 = ld4(row + 4 * ((8 * m + 2 * t + u) ^ sw));
              Up[m][u][0] = fma_f32x2(pack_f32x2(w4.x, w4.y), qp[d], Up[m][u][0]);
              Up[m][u][1] = fma_f32x2(pack_f32x2(w4.z, w4.w), qp[d], Up[m][u][1]);
            }
        }
#pragma unroll
        for (int m = 0; m < 8; ++m)
#pragma unroll
          for (int u = 0; u < 2; ++u) {
            unpack_f32x2(Up[m][u][0], Uf[m][u].x, Uf[m][u].y);
            unpack_f32x2(Up[m][u][1], Uf[m][u].z, Uf[m][u].w);
          }
      }
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[nt][k] = 0.f;
      const RegFrag A{Uf};
#pragma unroll
      for (int pass = 0; pass < 2; ++pass)
        mlp_pass16_dynamic(M, pass, g, t, lane, pik, p.pj_k, k_c, k_wimg, k_ln, A, acc[2 * pass], acc[2 * pass + 1]);
      // softmax over the 32 edges of heads g and g + 8 (see x2h_k_mma_kernel), times e_w
#pragma unroll
      for (int hs = 0; hs < 2; ++hs) {
        float l[8];
        float mx = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const bool valid = (vmask >> (8 * nt + 2 * t + v)) & 1u;
            l[2 * nt + v] = valid ? acc[nt][2 * hs + v] : -INFINITY;
            mx = fmaxf(mx, l[2 * nt + v]);
          }
        mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 1));
        mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 2));
        float sum = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) { l[k] = (mx == -INFINITY) ? 0.f : expf(l[k] - mx); sum += l[k]; }
        sum += __shfl_xor_sync(CBG_FULL, sum, 1);
        sum += __shfl_xor_sync(CBG_FULL, sum, 2);
        const float inv = (sum > 0.f) ? sum : 1.f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt)
#pragma unroll
          for (int v = 0; v < 2; ++v) {
            const int e = 8 * nt + 2 * t + v;
            alpha[e * CBG_HEADS + g + 8 * hs] = (l[2 * nt + v] / inv) * M.ew[e];
          }
      }
      pair_barrier(bar);                                            // (2) alpha ready
      pair_barrier(bar);                                            // (3) V warp done with M / alpha / next_slot
      if (p.ticket != nullptr && lane == 0) *next_slot = nxt;
    } else {
      // ---------------- V warp ----------------
      st4(piv + 4 * lane, ldg4(p.pi_v + (size_t)i * CBG_H + 4 * lane));
      pair_barrier(bar);                                            // (1) M ready (also orders the piv writes for this warp)
      float acc[4][4];
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int k = 0; k < 4; ++k) acc[nt][k] = 0.f;
      const SmemFrag A{v_w1, lane};
#pragma unroll
      for (int pass = 0; pass < 2; ++pass)
        mlp_pass16_dynamic(M, pass, g, t, lane, piv, p.pj_v, v_c, v_wimg, v_ln, A, acc[2 * pass], acc[2 * pass + 1]);
      pair_barrier(bar);                                            // (2) alpha ready
      // acc[nt][0..1]: head g, edges 8nt + 2t, +1; acc[nt][2..3]: head g + 8
      const float b1a = v_b1[g], b1b = v_b1[g + 8];
      float ax = 0.f, ay = 0.f, az = 0.f;
#pragma unroll
      for (int nt = 0; nt < 4; ++nt)
#pragma unroll
        for (int v = 0; v < 2; ++v) {
          const int e = 8 * nt + 2 * t + v;
          const float coef = alpha[e * CBG_HEADS + g] * (acc[nt][v] + b1a) + alpha[e * CBG_HEADS + g + 8] * (acc[nt][2 + v] + b1b);
          ax = fmaf(coef, M.rel[0][e], ax);
          ay = fmaf(coef, M.rel[1][e], ay);
          az = fmaf(coef, M.rel[2][e], az);
        }
      ax = warp_sum(ax); ay = warp_sum(ay); az = warp_sum(az);
      if (lane == 0) st4(p.dx + 4 * (size_t)n, make_float4(ax * (1.f / 16.f), ay * (1.f / 16.f), az * (1.f / 16.f), 0.f));
      pair_barrier(bar);                                            // (3) node done
    }
    if (p.ticket != nullptr) {
      pair_barrier(bar);                                            // (4) next node published
      n = *next_slot;
    } else {
      n += gridDim.x * kPairs;
    }
  }
}

int g_num_sms = 0;
int g_edge_warps = 12;
int g_edge_impl = 6;       // 6 (default): tcgen05 kernels (x2h_tc.cu); 0-5: the SIMT / mma.sync generations below (kept as tested alternatives)
int g_edge_mma_warps = 8;
int g_h2x_warps = 12;
int g_h2x_pairs = 4;       // node pairs per CTA of the pair kernel (4: 255 registers, 5: 204)
int g_h2x_impl = 0;        // 0 (default, measured faster at c2: 85 vs 91 us per launch): SIMT kernel; 1: tensor-core pair kernel

template <int W>
int set_attrs() {
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_k_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, x2hk_smem(W)));
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_v_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, x2hv_smem(W)));
  CBG_CUDA_OK(cudaFuncSetAttribute(h2x_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, h2x_smem(W)));
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_k_mma_kernel<W, (W > 8 ? 1 : 2)>, cudaFuncAttributeMaxDynamicSharedMemorySize, x2hk_mma_smem(W)));
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_v_mma_kernel<W>, cudaFuncAttributeMaxDynamicSharedMemorySize, x2hv_mma_smem(W)));
  if (W == 8) CBG_CUDA_OK(cudaFuncSetAttribute(h2x_pair_kernel<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, h2x_pair_smem(4)));
  if (W == 8) CBG_CUDA_OK(cudaFuncSetAttribute(h2x_pair_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, h2x_pair_smem(5)));
  if (W == 8) CBG_CUDA_OK(cudaFuncSetAttribute(x2h_k_mma2_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, x2hk_mma2_smem(8)));
  return 0;
}

int edge_grid(int n_nodes, int warps) {
  const int need = (n_nodes + warps - 1) / warps;
  return need < g_num_sms ? need : g_num_sms;
}

template <int W>
int launch_x2h(const EdgeArgs& a, cudaStream_t st) {
  const int grid = edge_grid(a.n_nodes, W);
  EdgeArgs av = a;
  if (av.ticket) av.ticket += 1;          // the second kernel has its own work counter
  CBG_PROF_BEGIN(CBG_K_X2H_K, st);
  x2h_k_kernel<W><<<grid, W * 32, x2hk_smem(W), st>>>(a);
  CBG_LAUNCHED(CBG_K_X2H_K, st);
  CBG_PROF_BEGIN(CBG_K_X2H_V, st);
  x2h_v_kernel<W><<<grid, W * 32, x2hv_smem(W), st>>>(av);
  CBG_LAUNCHED(CBG_K_X2H_V, st);
  return 0;
}

// impl 1: both kernels on the tensor cores; 2: tensor-core x2h_k + SIMT x2h_v; 3: SIMT x2h_k + tensor-core x2h_v;
// 4 / 5: x2h_k with the RBF mat-vec on the tensor cores as well (x2h_k_mma2_kernel) + SIMT / tensor-core x2h_v
// (the two halves agree on the layout of w, so they can be mixed: used by the tests to localise a mismatch)
template <int W, int WS>
int launch_x2h_mma(const EdgeArgs& a, cudaStream_t st, int impl) {
  EdgeArgs av = a;
  if (av.ticket) av.ticket += 1;          // the second kernel has its own work counter
  CBG_PROF_BEGIN(CBG_K_X2H_K, st);
  if (impl == 3) x2h_k_kernel<WS><<<edge_grid(a.n_nodes, WS), WS * 32, x2hk_smem(WS), st>>>(a);
  else if (impl >= 4) x2h_k_mma2_kernel<8><<<edge_grid(a.n_nodes, 8), 8 * 32, x2hk_mma2_smem(8), st>>>(a);
  else x2h_k_mma_kernel<W, (W > 8 ? 1 : 2)><<<edge_grid(a.n_nodes, W), W * 32, x2hk_mma_smem(W), st>>>(a);
  CBG_LAUNCHED(CBG_K_X2H_K, st);
  CBG_PROF_BEGIN(CBG_K_X2H_V, st);
  if (impl == 2 || impl == 4) x2h_v_kernel<WS><<<edge_grid(a.n_nodes, WS), WS * 32, x2hv_smem(WS), st>>>(av);
  else x2h_v_mma_kernel<W><<<edge_grid(a.n_nodes, W), W * 32, x2hv_mma_smem(W), st>>>(av);
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
  static bool done_dev[CBG_MAX_DEVICES] = {};
  bool& done = cbg_dev_flag(done_dev);
  if (done) return 0;
  int dev = 0;
  CBG_CUDA_OK(cudaGetDevice(&dev));
  CBG_CUDA_OK(cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev));
  if (const char* e = getenv("CBG_EDGE_WARPS")) {
    const int w = atoi(e);
    if (w == 8 || w == 12 || w == 16) g_edge_warps = w;
  }
  if (const char* e = getenv("CBG_EDGE_IMPL")) {
    if (strcmp(e, "simt") == 0) g_edge_impl = 0;
    else if (e[0] >= '0' && e[0] <= '6' && e[1] == 0) g_edge_impl = e[0] - '0';
  }
  if (const char* e = getenv("CBG_H2X_IMPL")) g_h2x_impl = (e[0] == '0') ? 0 : 1;
  if (const char* e = getenv("CBG_H2X_PAIRS")) g_h2x_pairs = (atoi(e) == 5) ? 5 : 4;
  if (const char* e = getenv("CBG_H2X_WARPS")) {
    const int w = atoi(e);
    if (w == 8 || w == 12 || w == 16) g_h2x_warps = w;
  }
  if (const char* e = getenv("CBG_EDGE_MMA_WARPS")) {
    const int w = atoi(e);
    if (w == 8 || w == 12 || w == 16) g_edge_mma_warps = w;
  }
  if (int rc = set_attrs<8>()) return rc;
  if (int rc = set_attrs<12>()) return rc;
  if (int rc = set_attrs<16>()) return rc;
  done = true;
  return 0;
}

int cbg_launch_rcache(const float* layers, int num_layers, const float4* x4, const int* snbr, int n_nodes,
                      float* rcache, cudaStream_t st) {
  if (n_nodes <= 0 || num_layers <= 0) return 0;
  if (int rc = cbg_edge_init()) return rc;
  static bool attr_dev[CBG_MAX_DEVICES] = {};
  bool& attr = cbg_dev_flag(attr_dev);
  if (!attr) {
    CBG_CUDA_OK(cudaFuncSetAttribute(rcache_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kRcSmem));
    attr = true;
  }
  dim3 grid((unsigned)((n_nodes + 7) / 8 < 4 * g_num_sms ? (n_nodes + 7) / 8 : 4 * g_num_sms), 2 * num_layers);
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  rcache_kernel<<<grid, 256, kRcSmem, st>>>(layers, cbg_layout::kLayerFloats, x4, snbr, n_nodes, rcache);
  CBG_LAUNCHED(CBG_K_MISC, st);
  return 0;
}

int cbg_launch_x2h(const EdgeArgs& a, cudaStream_t st) {
  if (a.n_nodes <= 0) return 0;
  if (int rc = cbg_edge_init()) return rc;
  if (g_edge_impl == 6) return cbg_launch_x2h_tc(a, st);       // tcgen05 kernels (x2h_tc.cu)
  if (g_edge_impl >= 1) {
    switch (g_edge_mma_warps) {
      case 12: return launch_x2h_mma<12, 12>(a, st, g_edge_impl);
      case 16: return launch_x2h_mma<16, 12>(a, st, g_edge_impl);
      default: return launch_x2h_mma<8, 12>(a, st, g_edge_impl);
    }
  }
  switch (g_edge_warps) {
    case 8: return launch_x2h<8>(a, st);
    case 16: return launch_x2h<16>(a, st);
    default: return launch_x2h<12>(a, st);
  }
}

int cbg_launch_h2x(const EdgeArgs& a, cudaStream_t st) {
  if (a.n_nodes <= 0) return 0;
  if (int rc = cbg_edge_init()) return rc;
  if (g_h2x_impl == 1) {
    const int kp = g_h2x_pairs;
    const int need = (a.n_nodes + kp - 1) / kp;
    const int grid = need < g_num_sms ? need : g_num_sms;
    CBG_PROF_BEGIN(CBG_K_H2X, st);
    if (kp == 5) h2x_pair_kernel<5><<<grid, 5 * 64, h2x_pair_smem(5), st>>>(a);
    else h2x_pair_kernel<4><<<grid, 4 * 64, h2x_pair_smem(4), st>>>(a);
    CBG_LAUNCHED(CBG_K_H2X, st);
    return 0;
  }
  switch (g_h2x_warps) {
    case 8: return launch_h2x<8>(a, st);
    case 16: return launch_h2x<16>(a, st);
    default: return launch_h2x<12>(a, st);
  }
}

// testing / tuning hook (include/cbg_b200.h): pick the X2H edge-kernel implementation and its warps per CTA
int cbg_edge_set_impl(int impl, int warps) {
  if (int rc = cbg_edge_init()) return rc;
  if (impl < 0 || impl > 6) { cbg_set_error("edge impl must be 0 (simt), 1 (mma), 2 (mma k + simt v), 3 (simt k + mma v), 4 (mma2 k + simt v), 5 (mma2 k + mma v) or 6 (tcgen05)"); return 1; }
  if (warps != 0 && warps != 8 && warps != 12 && warps != 16) { cbg_set_error("warps per CTA must be 8, 12 or 16"); return 1; }
  g_edge_impl = impl;
  if (warps) { if (impl) g_edge_mma_warps = warps; else g_edge_warps = warps; }
  return 0;
}

int cbg_edge_set_h2x_impl(int impl) {
  if (int rc = cbg_edge_init()) return rc;
  if (impl != 0 && impl != 1) { cbg_set_error("h2x impl must be 0 (simt) or 1 (tensor-core pair kernel)"); return 1; }
  g_h2x_impl = impl;
  return 0;
}
