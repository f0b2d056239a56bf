// X2H attention on the 5th-generation tensor cores: one CTA tile = 4 destination nodes x 32 in-edges = 128 edge rows.
//
// Reference semantics: repo/modules/attention/x2h_attention.py:43-97 (per edge e = (j -> i):
//   kv = [onehot(type) | onehot(type) (x) g(d) | h_i | h_j], k = MLP_k(kv), v = MLP_v(kv) * e_w,
//   alpha = softmax_j(<q_i, k_ij>/sqrt(8)) per head, h_i += sum_j alpha_ij v_ij).
//
// Both edge MLPs are 340 -> 128 -> LayerNorm -> ReLU -> 128.  Per 128-row tile the kernel runs two GEMMs on tcgen05:
//   MMA1  pre[128 x 128] = G[128 x 96] * Wg[96 x 128]      G = [onehot(t) (x) g(d) | onehot(t) | onehot(node slot)]
//                                                           Wg = [Wrf[t] ; c[t] ; Pi rows of the tile's 4 nodes]
//         (the type-dependent RBF mat-vec, the type bias and the destination-node plane Pi in one K = 96 product;
//          the source plane Pj[j] is added by the SIMT stage from rows staged in shared memory)
//   MMA2  out[128 x 128] = relu(LN(pre + Pj)) * W1^T
// fp32 accuracy comes from the split x = hi + lo into two f16 values (scaled by powers of two so lo stays normal)
// and the three products hi*hi + hi*lo + lo*hi accumulated in fp32 in TMEM (cbg_tc.cuh): same error class as 3xTF32
// at twice the tensor-core rate and half the operand bytes.
//
// A operands (G and the activations) live in TENSOR MEMORY (tcgen05.mma with A from TMEM: lane = edge row), written
// by tcgen05.st from the warps that produce them; B operands (weight images, pre-split and pre-laid-out by the host
// packer) stay resident in shared memory for the whole persistent CTA.  Nothing of size [E, 128] touches HBM and no
// R-cache is needed: the only per-edge gather is the 512-byte Pj row (cp.async, L2 resident).
//
// Warp roles (16 warps x 128 registers; the scheduler prefers higher warp ids, hence the order):
//   warps 0-3   EPI       thread = edge row, inputs prefetched one tile ahead.  k: <q_i, k> per head, softmax over the
//                         node's 32 edges through a shared-memory transpose, w = alpha * e_w;
//                         v: (v + b1v) * w, sum over the node's 32 edges, h_i += .
//   warps 4-11  S1        thread = (edge row, column half).  S1(tile t): TMEM(pre) + Pj -> LayerNorm (mean-free: the
//                         packer centres the first Linear over the feature axis) -> ReLU -> (hi, lo) f16 -> TMEM (in
//                         place).  Each thread also builds its half of the G row of tile t+1 (geometry, type, Gaussian
//                         smearing -> TMEM): values before, stores right after MMA1(t) has completed
//   warp 12     MMA       one lane issues every tcgen05.mma / commit (fully unrolled, tile-invariant descriptors)
//   warp 13     PROD-Pi   up to two tiles ahead: the Pi rows of the tile's four nodes into their K columns of the Wg images
//   warps 14-15 PROD-Pj   one tile ahead, two node slots each: cp.async of the node's 32 Pj rows into the slot's chunk
// Pipelining: TMEM holds two pre/activation buffers, so MMA1 of tile t+1 and MMA2 of tile t-1 run while S1 works on
// tile t and EPI on tile t-1.
//
// H2X (repo/modules/attention/h2x_attention.py:34-73) runs on the same kernel over the list of generated nodes:
//   launch 1  MODE_K with the xk / xq weights -> w = alpha * e_w (compact buffer, indexed by list position)
//   launch 2  MODE_XV with the xv weights: the second Linear has one output per head (N = 16 MMA), and the epilogue forms
//             dx_i = (1/16) sum_e (sum_hd w_e,hd (v_e,hd + b1_hd)) (x_i - x_j)     (mean over heads of alpha * v * e_w * rel_x)
// replacing the fp32 SIMT h2x_kernel (edge.cu), whose 4 GFLOP per launch ran at ~65 % of the fp32 pipe.
#include <math.h>
#include "cbg_kernels.cuh"
#include "cbg_tc.cuh"

using namespace cbg_tc;

namespace {

// ---- blob fields ------------------------------------------------------------------------------------------------
constexpr long long kOffKW1 = cbg_layout::layer_offset(CBG_LF_X2H_K_TCW1);
constexpr long long kOffKWg = cbg_layout::layer_offset(CBG_LF_X2H_K_TCWG);
constexpr long long kOffVW1 = cbg_layout::layer_offset(CBG_LF_X2H_V_TCW1);
constexpr long long kOffVWg = cbg_layout::layer_offset(CBG_LF_X2H_V_TCWG);
constexpr long long kOffKLn = cbg_layout::layer_offset(CBG_LF_X2H_K_LN);
constexpr long long kOffVLn = cbg_layout::layer_offset(CBG_LF_X2H_V_LN);
constexpr long long kOffVB1 = cbg_layout::layer_offset(CBG_LF_X2H_V_B1);
constexpr long long kOffRbf = cbg_layout::layer_offset(CBG_LF_X2H_K_RBF);
constexpr long long kOffXKW1 = cbg_layout::layer_offset(CBG_LF_H2X_K_TCW1);
constexpr long long kOffXKWg = cbg_layout::layer_offset(CBG_LF_H2X_K_TCWG);
constexpr long long kOffXVW1 = cbg_layout::layer_offset(CBG_LF_H2X_V_TCW1);
constexpr long long kOffXVWg = cbg_layout::layer_offset(CBG_LF_H2X_V_TCWG);
constexpr long long kOffXKLn = cbg_layout::layer_offset(CBG_LF_H2X_K_LN);
constexpr long long kOffXVLn = cbg_layout::layer_offset(CBG_LF_H2X_V_LN);
constexpr long long kOffXVB1 = cbg_layout::layer_offset(CBG_LF_H2X_V_B1);
constexpr long long kOffXRbf = cbg_layout::layer_offset(CBG_LF_H2X_RBF);

// what a launch computes; the weight fields it uses travel in TcWeights (filled by the launcher)
enum { MODE_K = 0, MODE_V = 1, MODE_XV = 2 };
struct TcWeights {
  const float* w1;     // (hi | lo) image of the second Linear: [128 n][128 k], MODE_XV [16 n][128 k]
  const float* wg;     // (hi | lo) image of [Wrf ; c ; Pi columns]
  const float* ln;     // gamma[128], beta[128]
  const float* b1;     // second-Linear bias (MODE_V: 128, MODE_XV: 16; unused by MODE_K - it cancels in the softmax)
  const float* rbf;    // Gaussian offsets [20], coefficient at [20]
};

// ---- scales (exact powers of two; must match modules.py: tc_f16_image) ----------------------------------------
constexpr float kScaleG = 1024.f;        // g(d) and the type one-hot in G
// Wrf, c in Wg are scaled by 16 (packer)                       -> pre accumulates at 2^14
constexpr float kInvPre = 1.f / 16384.f;
constexpr uint32_t kHalfTypeOne = 0x6400u;   // f16 1024
constexpr uint32_t kHalfNodeOne = 0x7400u;   // f16 16384: node one-hot x unscaled Pi = Pi * 2^14
constexpr float kScaleA = 64.f;          // activations
constexpr float kInvOut = 1.f / 4096.f;  // W1 image is scaled by 64 -> out accumulates at 2^12

// ---- shapes -----------------------------------------------------------------------------------------------------
constexpr int KG = 96;                   // K of MMA1 (84 used + 8 node one-hot columns (2 tile parities x 4) + 4 zero)
constexpr int KG_LO = 80;                // the lo part of G is non-zero only in the RBF columns
constexpr int NCH = 4;                   // Pj buffers: one 32-row chunk per node slot of a tile (refilled for tile t+1 as soon
                                         // as the slot's S1 warps have consumed tile t)
constexpr uint32_t PJ_ROW = 528;         // padded row stride: 16-byte row-per-lane reads are bank-conflict free
constexpr uint32_t PJ_CHUNK = 32 * PJ_ROW;
constexpr uint32_t W1_IMG = 128 * 128 * 2;            // one (hi | lo) image, bytes
constexpr uint32_t W1X_IMG = 16 * 128 * 2;            // MODE_XV: 16 output rows
constexpr uint32_t WG_IMG = 128 * KG * 2;
constexpr uint32_t W1_SBO = (128 / 8) * 128, WG_SBO = (KG / 8) * 128, LBO = 128;
constexpr uint32_t SM_W1 = 0;                         // hi | lo
constexpr uint32_t SM_WG = SM_W1 + 2 * W1_IMG;
constexpr uint32_t SM_PJ = SM_WG + 2 * WG_IMG;
constexpr uint32_t SM_LN = SM_PJ + NCH * PJ_CHUNK;    // gamma * 64 [128] | beta * 64 [128]
constexpr uint32_t SM_B1 = SM_LN + 1024;              // b1v [128]
constexpr uint32_t SM_RBF = SM_B1 + 512;              // Gaussian offsets [20] + coeff
constexpr uint32_t SM_XCH = SM_RBF + 128;              // sum-of-squares exchange between the two half-row S1 warps
constexpr uint32_t SM_QBUF = SM_XCH + 2048;             // EPI: [warp][tile parity][128] q row of the warp's node
constexpr uint32_t SM_SOFT = SM_QBUF + 4096;            // EPI: [warp][32 edges][17] logits <-> weights transpose
constexpr uint32_t SM_VRED = SM_QBUF;                   // EPI of the v kernel (aliases QBUF / SOFT): [warp][32 edges][36] transpose
constexpr uint32_t SM_JN = SM_QBUF + 4 * 32 * 36 * 4;   // S1: per-thread slot of the neighbour id prefetched two tiles ahead (cp.async)
constexpr uint32_t SM_BAR = SM_JN + 2 * 256 * 4;       // [0]: neighbour id, [256 + t]: node id of the tile three ahead
constexpr int NBAR = 13 + 2 * NCH;
constexpr uint32_t SM_TOTAL = SM_BAR + 8 * NBAR + 16;
static_assert(SM_TOTAL <= 232448, "shared memory budget");
enum { B_WFULL = 0, B_GREADY, B_UNUSED, B_ACC1 /*2*/ = 3, B_AREADY /*2*/ = 5, B_ACC2 /*2*/ = 7, B_ACC2FREE /*2*/ = 9,
       B_PJFULL = 11, B_PJFREE = 11 + NCH, B_PIREADY /*2*/ = 11 + 2 * NCH };
// TMEM columns
constexpr uint32_t TM_BUF = 0;           // 2 x 128: pre (fp32) -> a_hi (64 cols) | a_lo (64 cols)
constexpr uint32_t TM_OUT = 256;         // 128: output accumulator of MMA2
constexpr uint32_t TM_GHI = 384;         // 48 columns = 96 f16
constexpr uint32_t TM_GLO = 432;         // 40 columns = 80 f16
constexpr uint32_t TM_COLS = 512;
constexpr uint32_t IDESC128 = idesc_f16(128), IDESC16 = idesc_f16(16);

__device__ __forceinline__ int list_len(const EdgeArgs& p) {
  int n = p.n_nodes;
  if (p.n_nodes_dev) { const int nd = *p.n_nodes_dev; n = nd < n ? nd : n; }
  return n;
}
__device__ __forceinline__ int node_of(const EdgeArgs& p, int n, int n_list) {
  const int nc = n < n_list ? n : n_list - 1;
  return p.node_idx ? p.node_idx[nc] : nc;
}
__device__ __forceinline__ float warp_sum_x(float v) {      // fixed butterfly order: deterministic
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(CBG_FULL, v, o);
  return v;
}

// pipeline event stamps of CTA 0 (debugging; p.trace == nullptr in production: one predicated-off branch per event)
#define TC_STAMP(k, ev)                                                                                  \
  do {                                                                                                   \
    if (p.trace != nullptr && blockIdx.x == 0 && lane == 0 && (k) < p.trace_tiles) p.trace[(k) * 16 + (ev)] = clock64(); \
  } while (0)

// =================================================================================================================
template <int MODE>
__global__ void __launch_bounds__(512, 1) x2h_tc_kernel(EdgeArgs p, TcWeights W) {
  constexpr bool IS_V = MODE == MODE_V, IS_XV = MODE == MODE_XV;
  extern __shared__ __align__(1024) uint8_t smem[];
  const int n_list = list_len(p);
  const int n_tiles = (n_list + 3) >> 2;
  if ((int)blockIdx.x >= n_tiles) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bars = sbase + SM_BAR;
  auto bar = [&](int i) { return bars + 8u * (uint32_t)i; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 8 * NBAR);
  const int n_my = (n_tiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1;     // tiles of this CTA: blockIdx.x + k * gridDim.x
  // node of slot `slot` of this CTA's kk-th tile (clamped to the list: surplus slots of the last tile redo the last node)
  auto tile_node = [&](int kk, int slot) { return node_of(p, 4 * ((int)blockIdx.x + kk * (int)gridDim.x) + slot, n_list); };
  // row of the w buffer: the node id (X2H: [N, 32, 16]) or the list position (H2X: compact [n_list, 32, 16])
  auto w_row = [&](int kk, int slot, int i) {
    const int n = 4 * ((int)blockIdx.x + kk * (int)gridDim.x) + slot;
    return p.w_compact ? (n < n_list ? n : n_list - 1) : i;
  };

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TM_COLS);
  if (tid == 32) {
    mbar_init(bar(B_WFULL), 1);
    mbar_init(bar(B_GREADY), 8);
    mbar_init(bar(B_PIREADY), 1);
    mbar_init(bar(B_PIREADY + 1), 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(B_ACC1 + b), 1);
      mbar_init(bar(B_AREADY + b), 8);
      mbar_init(bar(B_ACC2 + b), 1);
      mbar_init(bar(B_ACC2FREE + b), 4);
    }
    for (int c = 0; c < NCH; ++c) { mbar_init(bar(B_PJFULL + c), 32); mbar_init(bar(B_PJFREE + c), 2); }
    fence_mbar_init();
  }
  {   // LayerNorm affine (pre-multiplied by the activation scale) and the value bias
    float* s_ln = reinterpret_cast<float*>(smem + SM_LN);
    float* s_b1 = reinterpret_cast<float*>(smem + SM_B1);
    if (tid < 256) s_ln[tid] = W.ln[tid] * kScaleA;
    else if (tid < 384) s_b1[tid - 256] = (IS_V || (IS_XV && tid - 256 < CBG_HEADS)) ? W.b1[tid - 256] : 0.f;
    else if (tid < 384 + 24) reinterpret_cast<float*>(smem + SM_RBF)[tid - 384] = W.rbf[tid - 384];
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp >= 4 && warp < 12) {
    // ===================================== S1 (tile k) + this thread's half of the G row of tile k + 1 ==================
    // thread = (edge row, column half hf).  S1: pre = TMEM + Pj -> LayerNorm -> ReLU -> (hi, lo) f16 -> TMEM.  The
    // first Linear is centred over the feature axis by the packer, so pre has zero mean and LayerNorm needs only the
    // sum of squares.  Right after MMA1(k) has completed (the wait below) the G region of TMEM may be rewritten: the
    // hf = 0 warps build the next tile's G rows there from coordinates prefetched one tile ahead.
    const int wq = warp & 3, hf = (warp >> 2) - 1;      // warps 4-7: column half 0, 8-11: half 1 (lane quarter = warp % 4)
    const uint32_t t_lane = tmem + ((uint32_t)(32 * wq) << 16);
    const float* s_ln = reinterpret_cast<const float*>(smem + SM_LN) + 64 * hf;
    float* s_x = reinterpret_cast<float*>(smem + SM_XCH);
    const int row = 32 * wq + lane;
    const float* rbf = reinterpret_cast<const float*>(smem + SM_RBF);
    const float c2 = rbf[20] * 1.4426950408889634f;      // exp(c u^2) = 2^(c log2(e) u^2)

    // G row of a tile, split between the two half-row warps of the quarter: this thread owns the Gaussians
    // m = 10 hf .. 10 hf + 9 of its edge row.  compute_g: values in registers (geometry, edge type, Gaussian smearing -
    // x2h_attention.py:46-52, unitransformer.py:88-99; the factor 1024 of the G scale rides in the exponent), BEFORE the
    // wait for MMA1 of the current tile; store_g: TMEM stores right after it, so MMA1 of the next tile can be issued as
    // early as possible.  G_hi: 48 columns (96 f16), G_lo: 40 columns; type block tb occupies columns 10 tb .. 10 tb + 9,
    // the type / node one-hots columns 40 .. 45 of G_hi (written by the half-0 warp).
    uint32_t ghi[5], glo[5];
    int t_e = 0;
    auto compute_g = [&](const float4 xi, const float4 xj) {
      const float rx = xi.x - xj.x, ry = xi.y - xj.y, rz = xi.z - xj.z;
      // explicit operation order: this lambda is inlined at two sites (tile 0 / the pipelined tiles) and the compiler
      // contracted x*x + y*y + z*z differently at each, so a node's result depended on its position in the CTA's tile list
      const float d = sqrtf(__fmaf_rn(rz, rz, __fmaf_rn(ry, ry, __fmul_rn(rx, rx))));
      const int fi = node_flags(xi), fj = node_flags(xj);
      t_e = ((fj & 1) ? 0 : 2) + ((fi & 1) ? 0 : 1);
#pragma unroll
      for (int mp = 0; mp < 5; ++mp) {
        const float u0 = d - rbf[10 * hf + 2 * mp], u1 = d - rbf[10 * hf + 2 * mp + 1];
        float g0, g1;
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g0) : "f"(fmaf(c2 * u0, u0, 10.f)));
        asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(g1) : "f"(fmaf(c2 * u1, u1, 10.f)));
        split_pair(g0, g1, ghi[mp], glo[mp]);
      }
    };
    auto store_g = [&](int kk) {
#pragma unroll
      for (int tb = 0; tb < 4; ++tb) {
        const bool on = t_e == tb;
        const uint32_t col = 10u * tb + 5u * (uint32_t)hf;
        tmem_st4(t_lane + TM_GHI + col, on ? ghi[0] : 0u, on ? ghi[1] : 0u, on ? ghi[2] : 0u, on ? ghi[3] : 0u);
        tmem_st1(t_lane + TM_GHI + col + 4u, on ? ghi[4] : 0u);
        tmem_st4(t_lane + TM_GLO + col, on ? glo[0] : 0u, on ? glo[1] : 0u, on ? glo[2] : 0u, on ? glo[3] : 0u);
        tmem_st1(t_lane + TM_GLO + col + 4u, on ? glo[4] : 0u);
      }
      if (hf == 0) {
        uint32_t w8[8];
#pragma unroll
        for (int cc = 0; cc < 8; ++cc) {
          const int col = 40 + cc;                   // f16 pair (2*col, 2*col + 1): k = 80 .. 95
          uint32_t val = 0u;
          if (col == 40) {
            val = (t_e == 0) ? kHalfTypeOne : ((t_e == 1) ? (kHalfTypeOne << 16) : 0u);
          } else if (col == 41) {
            val = (t_e == 2) ? kHalfTypeOne : ((t_e == 3) ? (kHalfTypeOne << 16) : 0u);
          } else if (col < 46) {
            const int kc = 84 + wq + 4 * (kk & 1);
            val = ((kc >> 1) == col) ? ((kc & 1) ? (kHalfNodeOne << 16) : kHalfNodeOne) : 0u;
          }
          w8[cc] = val;
        }
        tmem_st8(t_lane + TM_GHI + 40u, w8);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_GREADY));
    };
    // prefetch state of the G builders: coordinates of tile k+1 (C), node + neighbour of tile k+2 (B), node of tile k+3 (A).
    // The neighbour id of stage B travels through a per-thread shared-memory slot filled by a 4-byte cp.async: as a
    // register it was spilled right behind its LDG, and the spill store parked the warp on the load's latency once per
    // tile (ncu: long-scoreboard stalls on STL in the S1 loop).
    int iB = 0;
    float4 xiC = make_float4(0.f, 0.f, 0.f, 0.f), xjC = xiC;
    auto fetch_geo = [&](int i, int jn) { xiC = p.x4[i]; xjC = p.x4[jn >= 0 ? jn : i]; };
    const uint32_t s_jn = sbase + SM_JN + 4u * (uint32_t)(tid - 128);
    auto issue_jn = [&](int i) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_jn), "l"(p.nbr + (size_t)i * CBG_KMAX + lane) : "memory");
    };
    auto take_jn = [&]() {
      int v;
      asm volatile("cp.async.wait_all;\n\tld.shared.s32 %0, [%1];" : "=r"(v) : "r"(s_jn) : "memory");
      return v;
    };
    // node id of slot wq of this CTA's kk-th tile, the same way (stage A)
    auto issue_node = [&](int kk) {
      const int n = 4 * ((int)blockIdx.x + kk * (int)gridDim.x) + wq;
      const int nc = n < n_list ? n : n_list - 1;
      if (p.node_idx) asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s_jn + 1024u), "l"(p.node_idx + nc) : "memory");
      else asm volatile("st.shared.s32 [%0], %1;" ::"r"(s_jn + 1024u), "r"(nc) : "memory");
    };
    auto take_node = [&]() {
      int v;
      asm volatile("cp.async.wait_all;\n\tld.shared.s32 %0, [%1];" : "=r"(v) : "r"(s_jn + 1024u) : "memory");
      return v;
    };
    {   // tile 0 is produced here (its loads are exposed once per CTA); tiles 1.. through the prefetch pipeline
      const int i0 = tile_node(0, wq);
      fetch_geo(i0, p.nbr[(size_t)i0 * CBG_KMAX + lane]);
      compute_g(xiC, xjC);
      store_g(0);
      if (n_my > 1) {
        const int i1 = tile_node(1, wq);
        fetch_geo(i1, p.nbr[(size_t)i1 * CBG_KMAX + lane]);
      }
      if (n_my > 2) { iB = tile_node(2, wq); issue_jn(iB); }
      if (n_my > 3) issue_node(3);
    }
    for (int k = 0; k < n_my; ++k) {
      const int b = k & 1;
      const int c = wq;                                   // this quarter's Pj chunk, refilled once per tile
      if (k + 1 < n_my) compute_g(xiC, xjC);      // G values of tile k + 1 into registers
      mbar_wait(bar(B_ACC1 + b), (uint32_t)((k >> 1) & 1));      // MMA1(k) complete: pre is ready AND G may be rewritten
      tc_fence_after();
      if (warp == 4) TC_STAMP(k, 0);
      if (warp == 8) TC_STAMP(k, 5);
      if (k + 1 < n_my) store_g(k + 1);
      if (k + 2 < n_my) fetch_geo(iB, take_jn());  // coordinates of tile k + 2: in flight during the S1 body below
      if (warp == 4) TC_STAMP(k, 1);
      mbar_wait(bar(B_PJFULL + c), (uint32_t)(k & 1));
      // ---- S1
      const uint32_t t_buf = t_lane + TM_BUF + 128u * (uint32_t)b;
      float v[64];
      {
        uint32_t r[2][32];
        tmem_ld32_nowait(t_buf + 64u * hf, r[0]);
        tmem_ld32_nowait(t_buf + 64u * hf + 32u, r[1]);
        tmem_wait_ld();
        const uint8_t* prow = smem + SM_PJ + (uint32_t)c * PJ_CHUNK + (uint32_t)lane * PJ_ROW + 256u * hf;
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          const float4 pj = *reinterpret_cast<const float4*>(prow + 16 * j);
          const uint32_t* rv = &r[j >> 3][4 * (j & 7)];
          const float2 a0 = __ffma2_rn(make_float2(__uint_as_float(rv[0]), __uint_as_float(rv[1])),
                                       make_float2(kInvPre, kInvPre), make_float2(pj.x, pj.y));
          const float2 a1 = __ffma2_rn(make_float2(__uint_as_float(rv[2]), __uint_as_float(rv[3])),
                                       make_float2(kInvPre, kInvPre), make_float2(pj.z, pj.w));
          v[4 * j] = a0.x; v[4 * j + 1] = a0.y; v[4 * j + 2] = a1.x; v[4 * j + 3] = a1.y;
        }
      }
      if (warp == 4) TC_STAMP(k, 2);
      float2 q2 = make_float2(0.f, 0.f);
#pragma unroll
      for (int j = 0; j < 32; ++j) q2 = __ffma2_rn(make_float2(v[2 * j], v[2 * j + 1]), make_float2(v[2 * j], v[2 * j + 1]), q2);
      const float qs = q2.x + q2.y;
      s_x[256 * b + 128 * hf + row] = qs;
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PJFREE + c));          // this warp is done with the ring chunk
      asm volatile("bar.sync %0, 64;" ::"r"(1 + wq) : "memory");   // the two half-row warps of this row quarter
      if (warp == 4) TC_STAMP(k, 3);
      const float qo = s_x[256 * b + 128 * (hf ^ 1) + row];
      float rstd = rsqrtf((qs + qo) * (1.f / 128.f) + 1e-5f);      // MUFU.RSQ + one Newton step: < 1 ulp
      rstd = rstd * (1.5f - 0.5f * ((qs + qo) * (1.f / 128.f) + 1e-5f) * rstd * rstd);
      const float2 rr = make_float2(rstd, rstd);
      // relu((pre * rstd) * gamma + beta) * 64 -> (hi, lo) f16 into the buffer's columns: hi 0-63, lo 64-127
      // (the partner thread has read its accumulator columns before the barrier above)
#pragma unroll
      for (int ch = 0; ch < 2; ++ch) {
        uint32_t hi[16], lo[16];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 ga = *reinterpret_cast<const float4*>(s_ln + 32 * ch + 4 * j);
          const float4 be = *reinterpret_cast<const float4*>(s_ln + 128 + 32 * ch + 4 * j);
          const int e = 32 * ch + 4 * j;
          float2 y0 = __fmul2_rn(make_float2(v[e], v[e + 1]), rr);
          float2 y1 = __fmul2_rn(make_float2(v[e + 2], v[e + 3]), rr);
          y0 = __ffma2_rn(y0, make_float2(ga.x, ga.y), make_float2(be.x, be.y));
          y1 = __ffma2_rn(y1, make_float2(ga.z, ga.w), make_float2(be.z, be.w));
          split_pair_relu(y0.x, y0.y, hi[2 * j], lo[2 * j]);
          split_pair_relu(y1.x, y1.y, hi[2 * j + 1], lo[2 * j + 1]);
        }
        tmem_st16(t_buf + 32u * hf + 16u * ch, hi);
        tmem_st16(t_buf + 64u + 32u * hf + 16u * ch, lo);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_AREADY + b));
      // rotate the index prefetch after the register-pressure peak of the S1 body; consumed after the next tile's MMA1 wait
      if (k + 3 < n_my) { iB = take_node(); issue_jn(iB); }
      if (k + 4 < n_my) issue_node(k + 4);
      if (warp == 4) TC_STAMP(k, 4);
      if (warp == 8) TC_STAMP(k, 6);
    }
  } else if (warp < 4) {
    // ===================================== EPI (tile k), inputs prefetched one tile ahead ========================
    const int wq = warp;
    const uint32_t t_lane = tmem + ((uint32_t)(32 * wq) << 16);
    const float* s_b1 = reinterpret_cast<const float*>(smem + SM_B1);
    float* s_q = reinterpret_cast<float*>(smem + SM_QBUF) + wq * 256;          // [tile parity][128]: q row of the node
    float* s_sm = reinterpret_cast<float*>(smem + SM_SOFT) + wq * (32 * 17);   // [edge][17]: logits / weights transpose
    float* s_vr = reinterpret_cast<float*>(smem + SM_VRED) + wq * (32 * 36);   // v kernel: [edge][36] transpose (aliases the two above)
    // prefetched inputs of the next tile
    int i_n = tile_node(0, wq), i_nn = n_my > 1 ? tile_node(1, wq) : 0;
    int jn_n = -1;          // raw neighbour id of the next tile's edge (compared where it is used: no stall on the load)
    float ew_n = 0.f, h_n[4] = {0.f, 0.f, 0.f, 0.f};
    float4 w_n[4];
    auto prefetch = [&](int kk, int i) {
      const size_t eoff = (size_t)i * CBG_KMAX + lane;
      const size_t woff = ((size_t)w_row(kk, wq, i) * CBG_KMAX + lane) * CBG_HEADS;
      if constexpr (MODE == MODE_K) {
        jn_n = p.nbr[eoff];
        ew_n = p.ew[eoff];
        cp_async16(smem_u32(s_q + 128 * (kk & 1)) + 16u * (uint32_t)lane, p.q + (size_t)i * CBG_H + 4 * lane);
        asm volatile("cp.async.commit_group;" ::: "memory");
      } else {
        const float* wi = p.w + woff;
#pragma unroll
        for (int j = 0; j < 4; ++j) w_n[j] = ld4(wi + 4 * j);
        if constexpr (IS_V) {
#pragma unroll
          for (int j = 0; j < 4; ++j) h_n[j] = p.h[(size_t)i * CBG_H + 32 * j + lane];
        } else {      // MODE_XV: x_i - x_j of this lane's edge (padded slots: j = i, and their w is zero)
          const int jn = p.nbr[eoff];
          const float4 xi = p.x4[i], xj = p.x4[jn >= 0 ? jn : i];
          h_n[0] = xi.x - xj.x; h_n[1] = xi.y - xj.y; h_n[2] = xi.z - xj.z;
        }
      }
    };
    prefetch(0, i_n);
    for (int k = 0; k < n_my; ++k) {
      const int n = 4 * ((int)blockIdx.x + k * (int)gridDim.x) + wq;
      const bool live = n < n_list;
      const int i = i_n;
      const size_t eoff = (size_t)i * CBG_KMAX + lane;
      // take over this tile's inputs, start the next tile's
      const bool valid = jn_n >= 0;
      const float ew = ew_n;
      float wv[CBG_HEADS], hv[4];
      const size_t woff = ((size_t)w_row(k, wq, i) * CBG_KMAX + lane) * CBG_HEADS;
      if constexpr (MODE != MODE_K) {
#pragma unroll
        for (int j = 0; j < 4; ++j) { wv[4 * j] = w_n[j].x; wv[4 * j + 1] = w_n[j].y; wv[4 * j + 2] = w_n[j].z; wv[4 * j + 3] = w_n[j].w; hv[j] = h_n[j]; }
      } else {
        asm volatile("cp.async.wait_group 0;" ::: "memory");      // q row of this tile has landed (own copies only)
        __syncwarp();
      }
      i_n = i_nn;
      if (k + 1 < n_my) prefetch(k + 1, i_n);
      if (k + 2 < n_my) i_nn = tile_node(k + 2, wq);
      if constexpr (MODE == MODE_K) {
        const float* qs = s_q + 128 * (k & 1);
        float* my = s_sm + lane * 17;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 0) { mbar_wait(bar(B_ACC2), (uint32_t)(k & 1)); tc_fence_after(); }
          if (warp == 0) TC_STAMP(k, 7 + h);
          uint32_t r[2][32];
          tmem_ld32_nowait(t_lane + TM_OUT + 64u * h, r[0]);
          tmem_ld32_nowait(t_lane + TM_OUT + 64u * h + 32u, r[1]);
          tmem_wait_ld();
          if (h == 1) {      // the whole accumulator row is in registers: MMA2 of the next tile may overwrite it
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar(B_ACC2FREE));
          }
#pragma unroll
          for (int hh = 0; hh < 8; ++hh) {
            const float4 q0 = *reinterpret_cast<const float4*>(qs + 64 * h + 8 * hh);
            const float4 q1 = *reinterpret_cast<const float4*>(qs + 64 * h + 8 * hh + 4);
            const uint32_t* rv = &r[hh >> 2][8 * (hh & 3)];
            float2 tt = __fmul2_rn(make_float2(__uint_as_float(rv[0]), __uint_as_float(rv[1])), make_float2(q0.x, q0.y));
            tt = __ffma2_rn(make_float2(__uint_as_float(rv[2]), __uint_as_float(rv[3])), make_float2(q0.z, q0.w), tt);
            tt = __ffma2_rn(make_float2(__uint_as_float(rv[4]), __uint_as_float(rv[5])), make_float2(q1.x, q1.y), tt);
            tt = __ffma2_rn(make_float2(__uint_as_float(rv[6]), __uint_as_float(rv[7])), make_float2(q1.z, q1.w), tt);
            my[8 * h + hh] = valid ? (tt.x + tt.y) * kInvOut : -INFINITY;      // row = edge, stride 17: conflict free
          }
        }
        __syncwarp();
        // softmax over the 32 edges per head through the shared-memory transpose: lane = (head, half of the edges)
        {
          const int hd = lane & 15, e0 = 16 * (lane >> 4);
          float l[16];
          float mx = -INFINITY;
#pragma unroll
          for (int j = 0; j < 16; ++j) { l[j] = s_sm[(e0 + j) * 17 + hd]; mx = fmaxf(mx, l[j]); }
          mx = fmaxf(mx, __shfl_xor_sync(CBG_FULL, mx, 16));
          float sum = 0.f;
#pragma unroll
          for (int j = 0; j < 16; ++j) { l[j] = (mx == -INFINITY) ? 0.f : __expf(l[j] - mx); sum += l[j]; }
          sum += __shfl_xor_sync(CBG_FULL, sum, 16);
          const float inv = 1.f / ((sum > 0.f) ? sum : 1.f);
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 16; ++j) s_sm[(e0 + j) * 17 + hd] = l[j] * inv;
        }
        __syncwarp();
        if (live) {      // w = alpha * e_w, this lane's edge row
          const float sc = valid ? ew : 0.f;
          float* wo = p.w + woff;
#pragma unroll
          for (int j = 0; j < 4; ++j)
            st4(wo + 4 * j, make_float4(my[4 * j] * sc, my[4 * j + 1] * sc, my[4 * j + 2] * sc, my[4 * j + 3] * sc));
        }
        __syncwarp();
      } else if constexpr (IS_XV) {
        // H2X coordinate update: s_e = sum_hd w_e,hd (v_e,hd + b1_hd), dx_i = (1/16) sum_e s_e (x_i - x_j)
        mbar_wait(bar(B_ACC2), (uint32_t)(k & 1));
        tc_fence_after();
        if (warp == 0) TC_STAMP(k, 7);
        uint32_t r[16];
        tmem_ld16_nowait(t_lane + TM_OUT, r);
        tmem_wait_ld();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(bar(B_ACC2FREE));
        float sacc = 0.f;
#pragma unroll
        for (int hd = 0; hd < CBG_HEADS; ++hd) sacc = fmaf(fmaf(__uint_as_float(r[hd]), kInvOut, s_b1[hd]), wv[hd], sacc);
        const float ax = warp_sum_x(sacc * hv[0]), ay = warp_sum_x(sacc * hv[1]), az = warp_sum_x(sacc * hv[2]);
        if (live && lane == 0) st4(p.dx + 4 * (size_t)n, make_float4(ax * (1.f / 16.f), ay * (1.f / 16.f), az * (1.f / 16.f), 0.f));
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h == 0) { mbar_wait(bar(B_ACC2), (uint32_t)(k & 1)); tc_fence_after(); }
          if (warp == 0) TC_STAMP(k, 7 + h);
#pragma unroll
          for (int qq = 0; qq < 2; ++qq) {           // 32 columns at a time: features 64h + 32qq + (0..31)
            uint32_t r[32];
            tmem_ld32_nowait(t_lane + TM_OUT + 64u * h + 32u * qq, r);
            tmem_wait_ld();
            if (h == 1 && qq == 1) {      // last piece of the accumulator row is in registers
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive(bar(B_ACC2FREE));
            }
            // (v + b1v) * w -> this lane's row of the transpose buffer (stride 36 words: 16-byte stores conflict free),
            // then lane l sums column l over the node's 32 edges (fixed order: deterministic)
            float* vrow = s_vr + lane * 36;
#pragma unroll
            for (int c4 = 0; c4 < 8; ++c4) {
              const float4 b1 = *reinterpret_cast<const float4*>(s_b1 + 64 * h + 32 * qq + 4 * c4);
              const float wh = wv[8 * h + 4 * qq + (c4 >> 1)];
              float4 o;
              o.x = fmaf(__uint_as_float(r[4 * c4 + 0]), kInvOut, b1.x) * wh;
              o.y = fmaf(__uint_as_float(r[4 * c4 + 1]), kInvOut, b1.y) * wh;
              o.z = fmaf(__uint_as_float(r[4 * c4 + 2]), kInvOut, b1.z) * wh;
              o.w = fmaf(__uint_as_float(r[4 * c4 + 3]), kInvOut, b1.w) * wh;
              *reinterpret_cast<float4*>(vrow + 4 * c4) = o;
            }
            __syncwarp();
            float val[1];
            {
              float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
              for (int e = 0; e < 32; e += 4) {
                s0 += s_vr[(e + 0) * 36 + lane];
                s1 += s_vr[(e + 1) * 36 + lane];
                s2 += s_vr[(e + 2) * 36 + lane];
                s3 += s_vr[(e + 3) * 36 + lane];
              }
              val[0] = (s0 + s1) + (s2 + s3);
            }
            __syncwarp();
            if (live) p.h[(size_t)i * CBG_H + 64 * h + 32 * qq + lane] = hv[2 * h + qq] + val[0];
          }
        }
      }
      if (warp == 0) TC_STAMP(k, 9);
    }
  } else if (warp == 13) {
    // ===================================== PROD-Pi: the Pi rows of the tile's four nodes, up to two tiles ahead =======
    // The node's Pi row goes into its K column (84 + slot + 4 * tile parity) of the Wg images (hi, lo); the columns of a
    // tile parity were last read by MMA1 of tile - 2 (the wait below is always for the NEXT completion of that barrier,
    // so the parity wait is sound).  This warp never has copies in flight, which keeps its proxy fence cheap.
    const float* pi_plane = MODE != MODE_K ? p.pi_v : p.pi_k;
    float4 pi_c[4];
    int i_n[4];
#pragma unroll
    for (int sl = 0; sl < 4; ++sl) {
      pi_c[sl] = ldg4(pi_plane + (size_t)tile_node(0, sl) * CBG_H + 4 * lane);
      i_n[sl] = n_my > 1 ? tile_node(1, sl) : 0;
    }
    for (int kk = 0; kk < n_my; ++kk) {
      TC_STAMP(kk, 14);
      if (kk >= 2) mbar_wait(bar(B_ACC1 + (kk & 1)), (uint32_t)((((kk - 2) >> 1)) & 1));
      else mbar_wait(bar(B_WFULL), 0u);          // the columns live in the Wg images: the bulk copy must have landed
#pragma unroll
      for (int sl = 0; sl < 4; ++sl) {
        const float4 pi4 = pi_c[sl];
        if (kk + 1 < n_my) {      // next tile's row (its node id was fetched one tile earlier)
          pi_c[sl] = ldg4(pi_plane + (size_t)i_n[sl] * CBG_H + 4 * lane);
          if (kk + 2 < n_my) i_n[sl] = tile_node(kk + 2, sl);
        }
        const int kcol = 84 + sl + 4 * (kk & 1);
        const uint32_t cbase = (uint32_t)(kcol >> 3) * 128u + (uint32_t)(kcol & 7) * 2u;
        const float pv[4] = {pi4.x, pi4.y, pi4.z, pi4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int nn = 4 * lane + e;
          const uint32_t off = (uint32_t)(nn >> 3) * WG_SBO + (uint32_t)(nn & 7) * 16u + cbase;
          const __half hh = __float2half_rn(pv[e]);
          const __half hl = __float2half_rn(pv[e] - __half2float(hh));
          *reinterpret_cast<__half*>(smem + SM_WG + off) = hh;
          *reinterpret_cast<__half*>(smem + SM_WG + WG_IMG + off) = hl;
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar(B_PIREADY + (kk & 1)));
      TC_STAMP(kk, 15);
    }
  } else if (warp >= 14) {
    // ===================================== PROD-Pj: the Pj rows of two node slots, one tile ahead ====================
    // Warp 14 serves node slots 0 and 2, warp 15 slots 1 and 3: every chunk has ONE producer warp that walks the tiles in
    // order (a parity wait is only sound while the waiter can never be two phases ahead of the barrier).  Per (tile,
    // slot): cp.async of the node's 32 Pj rows into the slot's chunk as soon as the slot's S1 warps have consumed the
    // previous tile (warp = one row-coalesced 512-byte copy per instruction).
    const float* pj_plane = MODE != MODE_K ? p.pj_v : p.pj_k;
    const int s0 = warp - 14;                       // slots s0 and s0 + 2
    // neighbour ids are fetched one tile ahead and kept RAW (jn, i): the select jn >= 0 ? jn : i happens where the value is
    // consumed, one iteration later - selecting right after the load parked this warp on the load's latency (two
    // dependent global loads per slot) before it could serve the slot's copies (ncu: long-scoreboard stall at the select)
    int jn_c[2], ic_c[2], i_n[2];
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      ic_c[q] = tile_node(0, s0 + 2 * q);
      jn_c[q] = p.nbr[(size_t)ic_c[q] * CBG_KMAX + lane];
      i_n[q] = n_my > 1 ? tile_node(1, s0 + 2 * q) : 0;
    }
    for (int kk = 0; kk < n_my; ++kk) {
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        const int slot = s0 + 2 * q;
        const int jj = jn_c[q] >= 0 ? jn_c[q] : ic_c[q];
        if (kk + 1 < n_my) {      // next tile's neighbours (its node id was fetched one tile earlier)
          ic_c[q] = i_n[q];
          jn_c[q] = p.nbr[(size_t)i_n[q] * CBG_KMAX + lane];
          if (kk + 2 < n_my) i_n[q] = tile_node(kk + 2, slot);
        }
        if (kk >= 1) mbar_wait(bar(B_PJFREE + slot), (uint32_t)((kk - 1) & 1));
        const uint32_t dst = sbase + SM_PJ + (uint32_t)slot * PJ_CHUNK + 16u * (uint32_t)lane;
        int jr[32];
#pragma unroll
        for (int r = 0; r < 32; ++r) jr[r] = __shfl_sync(CBG_FULL, jj, r);       // all shuffles first: no per-row latency chain
#pragma unroll
        for (int r = 0; r < 32; ++r) cp_async16(dst + (uint32_t)r * PJ_ROW, pj_plane + (size_t)jr[r] * CBG_H + 4 * lane);
        cp_async_arrive(bar(B_PJFULL + slot));
      }
    }
  } else if (warp == 12) {
    // ===================================== MMA issuer ============================================================
    if (lane == 0) {
      constexpr uint32_t W1B = IS_XV ? W1X_IMG : W1_IMG;      // bytes of one second-Linear image
      constexpr uint32_t IDESC2 = IS_XV ? IDESC16 : IDESC128;
      mbar_expect_tx(bar(B_WFULL), 2 * W1B + 2 * WG_IMG);
      bulk_g2s(sbase + SM_W1, W.w1, 2 * W1B, bar(B_WFULL));
      bulk_g2s(sbase + SM_WG, W.wg, 2 * WG_IMG, bar(B_WFULL));
      mbar_wait(bar(B_WFULL), 0u);
      // Descriptors are tile-invariant: build the four bases once; a K step of 16 f16 (two core matrices, 256 bytes)
      // adds 16 to the 14-bit start-address field, so every MMA below costs one integer add and the issue itself
      // (the loops are fully unrolled - a rolled loop spends ~100 cycles per MMA on the uniform datapath, which made the
      // single issuing thread, not the tensor pipe, the limiter of the whole kernel).
      const uint64_t dg_hi = smem_desc(sbase + SM_WG, LBO, WG_SBO), dg_lo = smem_desc(sbase + SM_WG + WG_IMG, LBO, WG_SBO);
      const uint64_t d1_hi = smem_desc(sbase + SM_W1, LBO, W1_SBO), d1_lo = smem_desc(sbase + SM_W1 + W1B, LBO, W1_SBO);
      auto issue_mma2 = [&](int kk) {
        const int bb = kk & 1;
        mbar_wait(bar(B_AREADY + bb), (uint32_t)((kk >> 1) & 1));
        tc_fence_after();
        TC_STAMP(kk, 12);
        const uint32_t a_hi = tmem + TM_BUF + 128u * (uint32_t)bb, a_lo = a_hi + 64u;
        // one N = 128 accumulator: an MMA costs ~64 cycles whether N is 64 or 128 (measured: two N = 64 halves took 3.0K
        // cycles per tile, twice the N = 128 figure), so the output is not split
        if (kk > 0) { mbar_wait(bar(B_ACC2FREE), (uint32_t)((kk - 1) & 1)); tc_fence_after(); }
        const uint32_t d = tmem + TM_OUT;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks)     // small terms first
          umma_f16_ts(d, a_lo + 8u * ks, d1_hi + 16u * ks, IDESC2, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_f16_ts(d, a_hi + 8u * ks, d1_lo + 16u * ks, IDESC2, 1u);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) umma_f16_ts(d, a_hi + 8u * ks, d1_hi + 16u * ks, IDESC2, 1u);
        umma_commit(bar(B_ACC2));
        TC_STAMP(kk, 13);
      };
      for (int k = 0; k < n_my; ++k) {
        const int b = k & 1;
        mbar_wait(bar(B_GREADY), (uint32_t)(k & 1));
        mbar_wait(bar(B_PIREADY + b), (uint32_t)((k >> 1) & 1));
        tc_fence_after();
        TC_STAMP(k, 10);
        const uint32_t d = tmem + TM_BUF + 128u * (uint32_t)b;
#pragma unroll
        for (int ks = 0; ks < KG_LO / 16; ++ks)
          umma_f16_ts(d, tmem + TM_GLO + 8u * ks, dg_hi + 16u * ks, IDESC128, ks > 0 ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KG / 16; ++ks) umma_f16_ts(d, tmem + TM_GHI + 8u * ks, dg_lo + 16u * ks, IDESC128, 1u);
#pragma unroll
        for (int ks = 0; ks < KG / 16; ++ks) umma_f16_ts(d, tmem + TM_GHI + 8u * ks, dg_hi + 16u * ks, IDESC128, 1u);
        umma_commit(bar(B_ACC1 + b));
        TC_STAMP(k, 11);
        if (k > 0) issue_mma2(k - 1);
      }
      issue_mma2(n_my - 1);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TM_COLS);
}

// =================================================================================================================
// Hardware self-test of the operand conventions the kernel above relies on (A from TMEM with two K-consecutive f16
// per column, B in the canonical K-major no-swizzle layout, fp32 accumulator read back with tcgen05.ld 32x32b).
// D[128 x 128] = A[128 x 32] * B[128 x 32]^T, one CTA of 128 threads.  a, b: f16 row-major [128][32]; d: fp32 [128][128].
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const __half* a, const __half* b, float* d, int a_from_smem) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr uint32_t SBO = (32 / 8) * 128;        // K = 32: 4 core matrices per 8-row group
  const uint32_t sbase = smem_u32(smem);
  const uint32_t sb_b = sbase, sb_a = sbase + 128 * 32 * 2, bar0 = sbase + 2 * 128 * 32 * 2;
  uint32_t* slot = reinterpret_cast<uint32_t*>(smem + 2 * 128 * 32 * 2 + 8);
  if (warp == 0) tmem_alloc(smem_u32(slot), 256);
  if (tid == 32) { mbar_init(bar0, 1); fence_mbar_init(); }
  for (int e = tid; e < 128 * 32; e += 128) {
    const int r = e >> 5, kk = e & 31;
    const uint32_t off = (uint32_t)(r >> 3) * SBO + (uint32_t)(kk >> 3) * 128u + (uint32_t)(r & 7) * 16u + (uint32_t)(kk & 7) * 2u;
    *reinterpret_cast<__half*>(smem + off) = b[e];
    *reinterpret_cast<__half*>(smem + 128 * 32 * 2 + off) = a[e];
  }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  const uint32_t t_lane = tmem + ((uint32_t)(32 * warp) << 16);
  if (!a_from_smem) {     // row tid of A -> TMEM columns 128..143 (16 columns = 32 f16)
    uint32_t w16[16];
    const uint32_t* arow = reinterpret_cast<const uint32_t*>(a + (size_t)tid * 32);
#pragma unroll
    for (int c = 0; c < 16; ++c) w16[c] = arow[c];
    tmem_st16(t_lane + 128u, w16);
    tmem_wait_st();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  if (tid == 0) {
    for (int ks = 0; ks < 2; ++ks) {
      if (a_from_smem) umma_f16_ss(tmem, smem_desc(sb_a + 256u * ks, 128, SBO), smem_desc(sb_b + 256u * ks, 128, SBO), idesc_f16(128), ks > 0);
      else umma_f16_ts(tmem, tmem + 128u + 8u * ks, smem_desc(sb_b + 256u * ks, 128, SBO), idesc_f16(128), ks > 0);
    }
    umma_commit(bar0);
  }
  mbar_wait(bar0, 0u);
  tc_fence_after();
#pragma unroll 1
  for (int q = 0; q < 4; ++q) {
    uint32_t r[32];
    tmem_ld32_nowait(t_lane + 32u * q, r);
    tmem_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; ++j) d[(size_t)tid * 128 + 32 * q + j] = __uint_as_float(r[j]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, 256);
}

int g_tc_sms = 0;
long long* g_tc_trace = nullptr;
int g_tc_trace_tiles = 0;

int tc_init() {
  static bool done_dev[CBG_MAX_DEVICES] = {};
  bool& done = cbg_dev_flag(done_dev);
  if (done) return 0;
  int dev = 0;
  CBG_CUDA_OK(cudaGetDevice(&dev));
  CBG_CUDA_OK(cudaDeviceGetAttribute(&g_tc_sms, cudaDevAttrMultiProcessorCount, dev));
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_tc_kernel<MODE_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_tc_kernel<MODE_V>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
  CBG_CUDA_OK(cudaFuncSetAttribute(x2h_tc_kernel<MODE_XV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
  done = true;
  return 0;
}

}  // namespace

int cbg_launch_x2h_tc(const EdgeArgs& a, cudaStream_t st) {
  if (a.n_nodes <= 0) return 0;
  if (int rc = tc_init()) return rc;
  const int tiles = (a.n_nodes + 3) / 4;
  const int grid = tiles < g_tc_sms ? tiles : g_tc_sms;
  EdgeArgs ak = a;
  ak.w_compact = 0;
  ak.trace = g_tc_trace; ak.trace_tiles = g_tc_trace_tiles;          // debugging hook: one-shot, the next k kernel launch only
  g_tc_trace = nullptr;
  EdgeArgs av = a;
  av.w_compact = 0;
  const float* L = a.layer;
  const TcWeights wk{L + kOffKW1, L + kOffKWg, L + kOffKLn, nullptr, L + kOffRbf};
  const TcWeights wv{L + kOffVW1, L + kOffVWg, L + kOffVLn, L + kOffVB1, L + kOffRbf};
  CBG_PROF_BEGIN(CBG_K_X2H_K, st);
  x2h_tc_kernel<MODE_K><<<grid, 512, SM_TOTAL, st>>>(ak, wk);
  CBG_LAUNCHED(CBG_K_X2H_K, st);
  CBG_PROF_BEGIN(CBG_K_X2H_V, st);
  x2h_tc_kernel<MODE_V><<<grid, 512, SM_TOTAL, st>>>(av, wv);
  CBG_LAUNCHED(CBG_K_X2H_V, st);
  return 0;
}

// H2X for the listed (generated) nodes: attention weights with the xk / xq weights into the compact buffer a.w
// ([n_nodes, 32, 16]), then the value head + coordinate update into a.dx ([n_nodes, 4])
int cbg_launch_h2x_tc(const EdgeArgs& a, cudaStream_t st) {
  if (a.n_nodes <= 0) return 0;
  if (int rc = tc_init()) return rc;
  if (a.node_idx == nullptr || a.w == nullptr || a.dx == nullptr) { cbg_set_error("h2x_tc: node list, w and dx buffers are required"); return 1; }
  const int tiles = (a.n_nodes + 3) / 4;
  const int grid = tiles < g_tc_sms ? tiles : g_tc_sms;
  EdgeArgs ax = a;
  ax.w_compact = 1; ax.trace = nullptr; ax.trace_tiles = 0;
  const float* L = a.layer;
  const TcWeights wk{L + kOffXKW1, L + kOffXKWg, L + kOffXKLn, nullptr, L + kOffXRbf};
  const TcWeights wv{L + kOffXVW1, L + kOffXVWg, L + kOffXVLn, L + kOffXVB1, L + kOffXRbf};
  CBG_PROF_BEGIN(CBG_K_H2X, st);
  x2h_tc_kernel<MODE_K><<<grid, 512, SM_TOTAL, st>>>(ax, wk);
  CBG_LAUNCHED(CBG_K_H2X, st);
  CBG_PROF_BEGIN(CBG_K_H2X, st);
  x2h_tc_kernel<MODE_XV><<<grid, 512, SM_TOTAL, st>>>(ax, wv);
  CBG_LAUNCHED(CBG_K_H2X, st);
  return 0;
}

void cbg_x2h_tc_set_trace(long long* buf, int max_tiles) { g_tc_trace = buf; g_tc_trace_tiles = buf ? max_tiles : 0; }

int cbg_launch_umma_selftest(const void* a, const void* b, float* d, int a_from_smem, cudaStream_t st) {
  const int smem_bytes = 2 * 128 * 32 * 2 + 64;
  umma_selftest_kernel<<<1, 128, smem_bytes, st>>>((const __half*)a, (const __half*)b, d, a_from_smem);
  CBG_CUDA_OK(cudaGetLastError());
  return 0;
}
