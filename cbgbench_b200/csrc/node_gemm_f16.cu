// Node-level projections on tcgen05 with the f16 (hi, lo) split (kind::f16, three products per K step, fp32 accumulate
// in TMEM): same contract as node_gemm_tc.cu (planes Pj_k, Pj_v, Pi_k, Pi_v, q of one attention sub-layer; reference:
// x2h_attention.py:58-83, h2x_attention.py:42-62, common.py:151-171), half the tensor-core time and half the operand
// bytes of the 3xTF32 version at the same accuracy class (cbg_tc.cuh).
//
// One CTA = one 128-row tile.  A (rows of h, scaled by 16) is split once per CTA into hi/lo f16 tiles in shared memory
// (UMMA canonical K-major layout); the weight planes (scaled by 256) are pre-split and pre-laid-out by the packer in
// 64-wide K chunks (hi | lo = 32 KB) and stream through a 3-stage ring with cp.async.bulk + mbarrier; warp-specialised:
// 8 staging/epilogue warps, a copy thread, an MMA thread, two TMEM accumulators (plane g+1 accumulates while plane g
// drains).  LayerNorm + ReLU of the q MLP is done out of TMEM and fed back as the A operand of its second Linear.
#include "cbg_kernels.cuh"
#include "cbg_tc.cuh"

using namespace cbg_tc;

namespace {

constexpr int TM = 128;                        // rows per CTA (UMMA M)
constexpr int KC = 64;                         // K elements per weight chunk
constexpr int NKC = CBG_H / KC;                // 2 chunks per plane
constexpr int STAGES = 3;
constexpr uint32_t A_TILE = TM * CBG_H * 2;                // 32 KB per (hi | lo)
constexpr uint32_t B_CHUNK = 128 * KC * 2;                 // 16 KB per (hi | lo)
constexpr uint32_t B_STAGE = 2 * B_CHUNK;
constexpr uint32_t SM_A_HI = 0, SM_A_LO = A_TILE, SM_B0 = 2 * A_TILE;
constexpr uint32_t SM_BARS = SM_B0 + STAGES * B_STAGE;
constexpr uint32_t SM_TOTAL = SM_BARS + 128;
constexpr uint32_t A_SBO = (CBG_H / 8) * 128, B_SBO = (KC / 8) * 128, LBO = 128;
constexpr uint32_t TMEM_COLS = 256;
constexpr uint32_t IDESC = idesc_f16(128);
constexpr float kScaleA = 16.f;                // h (and the LayerNorm'ed q hidden) in the A tiles
constexpr float kInvAcc = 1.f / 4096.f;        // weights are scaled by 256 (packer): accumulator at 2^12

// byte offset of elements (row, k .. k+7) inside an A tile (k a multiple of 8: one 16-byte core-matrix row)
__device__ __forceinline__ uint32_t a_off8(int row, int k8) {
  return (uint32_t)(row >> 3) * A_SBO + (uint32_t)k8 * 128u + (uint32_t)(row & 7) * 16u;
}
__device__ __forceinline__ void store_split8(uint8_t* smem, int row, int k8, const float4 a, const float4 b) {
  uint4 hi, lo;
  split_pair(a.x * kScaleA, a.y * kScaleA, hi.x, lo.x);
  split_pair(a.z * kScaleA, a.w * kScaleA, hi.y, lo.y);
  split_pair(b.x * kScaleA, b.y * kScaleA, hi.z, lo.z);
  split_pair(b.z * kScaleA, b.w * kScaleA, hi.w, lo.w);
  const uint32_t off = a_off8(row, k8);
  *reinterpret_cast<uint4*>(smem + SM_A_HI + off) = hi;
  *reinterpret_cast<uint4*>(smem + SM_A_LO + off) = lo;
}
// weight chunk i lives at: plane(i / NKC) -> image index, chunk (i % NKC); image = NKC x (hi | lo) x [128 n][64 k] f16
__device__ __forceinline__ const float* chunk_src(const NodeGemmArgs& p, int i) {
  const int g = i / NKC, c = i % NKC;
  const int plane = (g < p.n_planes) ? (p.tc_first_plane + g) : 5;       // plane 5 = q second Linear
  return p.tch_planes + (size_t)plane * (NKC * B_STAGE / 4) + (size_t)c * (B_STAGE / 4);
}

__global__ void __launch_bounds__(320, 1) node_gemm_f16_kernel(NodeGemmArgs p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  if (p.n_rows_dev) {
    const int nd = *p.n_rows_dev;
    p.n_rows = nd < p.n_rows ? nd : p.n_rows;
  }
  const int row0 = blockIdx.x * TM;
  if (row0 >= p.n_rows) return;
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar_full = sbase + SM_BARS;                  // [STAGES]
  const uint32_t bar_empty = bar_full + 8 * STAGES;           // [STAGES]
  const uint32_t bar_acc_full = bar_empty + 8 * STAGES;       // [2]
  const uint32_t bar_acc_free = bar_acc_full + 16;            // [2]
  const uint32_t bar_a_ready = bar_acc_free + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BARS + 8 * (2 * STAGES + 5));

  if (warp == 0) tmem_alloc(smem_u32(tmem_slot), TMEM_COLS);
  if (tid == 32) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(bar_acc_full + 8 * b, 1); mbar_init(bar_acc_free + 8 * b, 8); }
    mbar_init(bar_a_ready, 8);
    fence_mbar_init();
  }
  const int n_gemm = p.n_planes + (p.has_q ? 1 : 0);
  const int total_chunks = n_gemm * NKC;

  if (warp < 8) {
    // A tile: rows of h -> (hi, lo) f16 tiles; lane <-> row keeps the 16-byte shared stores conflict free
    const int r = tid & (TM - 1);
    const int row = row0 + r;
    const bool live = row < p.n_rows;
    const float* arow = p.a + (size_t)(live ? (p.row_idx ? p.row_idx[row] : row) : 0) * CBG_H;
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int it = 0; it < 8; it += 4) {            // thread handles k8 = (tid >> 7) + 2 * j, j < 8
      float4 v[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k8 = (tid >> 7) + 2 * (it + j);
        v[2 * j] = live ? ldg4(arow + 8 * k8) : z;
        v[2 * j + 1] = live ? ldg4(arow + 8 * k8 + 4) : z;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) store_split8(smem, r, (tid >> 7) + 2 * (it + j), v[2 * j], v[2 * j + 1]);
    }
    fence_proxy_async();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 9) {
    // ===== weight-chunk producer =====
    if (lane == 0) {
      for (int i = 0; i < total_chunks; ++i) {
        const int s = i % STAGES;
        if (i >= STAGES) mbar_wait(bar_empty + 8 * s, (uint32_t)(((i / STAGES) - 1) & 1));
        mbar_expect_tx(bar_full + 8 * s, B_STAGE);
        bulk_g2s(sbase + SM_B0 + s * B_STAGE, chunk_src(p, i), B_STAGE, bar_full + 8 * s);
      }
    }
  } else if (warp == 8) {
    // ===== MMA issuer =====
    if (lane == 0) {
      const uint64_t da_hi = smem_desc(sbase + SM_A_HI, LBO, A_SBO), da_lo = smem_desc(sbase + SM_A_LO, LBO, A_SBO);
      for (int g = 0; g < n_gemm; ++g) {
        const int buf = g & 1;
        if (g >= 2) mbar_wait(bar_acc_free + 8 * buf, (uint32_t)(((g >> 1) - 1) & 1));   // epilogue drained this accumulator
        if (p.has_q && g == p.n_planes) mbar_wait(bar_a_ready, 0u);                      // A tiles now hold relu(LN(q_hidden))
        tc_fence_after();
        const uint32_t d_tmem = tmem + (uint32_t)(buf * 128);
#pragma unroll
        for (int c = 0; c < NKC; ++c) {
          const int i = g * NKC + c, s = i % STAGES;
          mbar_wait(bar_full + 8 * s, (uint32_t)((i / STAGES) & 1));
          tc_fence_after();
          const uint64_t db_hi = smem_desc(sbase + SM_B0 + s * B_STAGE, LBO, B_SBO);
          const uint64_t db_lo = smem_desc(sbase + SM_B0 + s * B_STAGE + B_CHUNK, LBO, B_SBO);
#pragma unroll
          for (int ks = 0; ks < KC / 16; ++ks) {
            const uint64_t ka = (uint64_t)(16 * (c * (KC / 16) + ks)), kb = (uint64_t)(16 * ks);   // 256 bytes per K step
            umma_f16_ss(d_tmem, da_lo + ka, db_hi + kb, IDESC, (c == 0 && ks == 0) ? 0u : 1u);    // small terms first
            umma_f16_ss(d_tmem, da_hi + ka, db_lo + kb, IDESC, 1u);
            umma_f16_ss(d_tmem, da_hi + ka, db_hi + kb, IDESC, 1u);
          }
          umma_commit(bar_empty + 8 * s);
        }
        umma_commit(bar_acc_full + 8 * buf);
      }
    }
  } else {
    // ===== epilogue warps =====
    const int q4 = warp & 3, chalf = warp >> 2;
    const int my_row = 32 * q4 + lane;
    const int grow = row0 + my_row;
    const int dst = (grow < p.n_rows) ? (p.row_idx ? p.row_idx[grow] : grow) : -1;
    for (int g = 0; g < n_gemm; ++g) {
      const int buf = g & 1;
      mbar_wait(bar_acc_full + 8 * buf, (uint32_t)((g >> 1) & 1));
      tc_fence_after();
      const uint32_t t_lane = tmem + ((uint32_t)(32 * q4) << 16) + (uint32_t)(buf * 128);
      const bool is_qhid = p.has_q && (g == p.n_planes - 1);
      const bool is_q2 = p.has_q && (g == p.n_planes);
      if (!is_qhid) {
        const float* bias = is_q2 ? p.q_b1 : (p.bias + g * CBG_H);
        float* out = is_q2 ? p.out_q : p.out[g];
#pragma unroll 1
        for (int cb = 0; cb < 2; ++cb) {
          const int col0 = chalf * 64 + cb * 32;
          uint32_t r[32];
          tmem_ld32_nowait(t_lane + (uint32_t)col0, r);
          tmem_wait_ld();
          if (dst >= 0) {
            float* o = out + (size_t)dst * CBG_H + col0;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const float4 b = ldg4(bias + col0 + 4 * j);
              st4(o + 4 * j, make_float4(fmaf(__uint_as_float(r[4 * j]), kInvAcc, b.x), fmaf(__uint_as_float(r[4 * j + 1]), kInvAcc, b.y),
                                         fmaf(__uint_as_float(r[4 * j + 2]), kInvAcc, b.z), fmaf(__uint_as_float(r[4 * j + 3]), kInvAcc, b.w)));
            }
          }
        }
      } else {
        if (chalf == 0) {
          // q hidden: + bias, two-pass LayerNorm over the row (TMEM is re-read instead of keeping 128 values live),
          // ReLU, back into the A tiles as the operand of the second Linear
          const float* bias = p.bias + g * CBG_H;
          float s = 0.f;
#pragma unroll 1
          for (int cb = 0; cb < 4; ++cb) {
            uint32_t r[32];
            tmem_ld32_nowait(t_lane + (uint32_t)(cb * 32), r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) s += fmaf(__uint_as_float(r[j]), kInvAcc, __ldg(bias + cb * 32 + j));
          }
          const float mean = s * (1.f / 128.f);
          float q = 0.f;
#pragma unroll 1
          for (int cb = 0; cb < 4; ++cb) {
            uint32_t r[32];
            tmem_ld32_nowait(t_lane + (uint32_t)(cb * 32), r);
            tmem_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) { const float d = fmaf(__uint_as_float(r[j]), kInvAcc, __ldg(bias + cb * 32 + j)) - mean; q = fmaf(d, d, q); }
          }
          const float rstd = 1.f / sqrtf(q * (1.f / 128.f) + 1e-5f);
#pragma unroll 1
          for (int cb = 0; cb < 4; ++cb) {
            uint32_t r[32];
            tmem_ld32_nowait(t_lane + (uint32_t)(cb * 32), r);
            tmem_wait_ld();
#pragma unroll
            for (int k8 = 0; k8 < 4; ++k8) {
              float a[8];
#pragma unroll
              for (int e = 0; e < 8; ++e) {
                const int col = cb * 32 + 8 * k8 + e;
                const float x = fmaf(__uint_as_float(r[8 * k8 + e]), kInvAcc, __ldg(bias + col)) - mean;
                a[e] = fmaxf(fmaf(x * rstd, __ldg(p.q_ln + col), __ldg(p.q_ln + 128 + col)), 0.f);
              }
              store_split8(smem, my_row, cb * 4 + k8, make_float4(a[0], a[1], a[2], a[3]), make_float4(a[4], a[5], a[6], a[7]));
            }
          }
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_a_ready);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_acc_free + 8 * buf);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

int cbg_launch_node_gemm_f16(const NodeGemmArgs& a, cudaStream_t st) {
  if (a.n_rows <= 0) return 0;
  if (!a.tch_planes) { cbg_set_error("f16 tensor-core node GEMM needs the f16 weight images (tch_planes)"); return 1; }
  static bool attr_dev[CBG_MAX_DEVICES] = {};
  bool& attr_set = cbg_dev_flag(attr_dev);
  if (!attr_set) {
    CBG_CUDA_OK(cudaFuncSetAttribute(node_gemm_f16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL));
    attr_set = true;
  }
  CBG_PROF_BEGIN(CBG_K_NODE_GEMM, st);
  node_gemm_f16_kernel<<<(a.n_rows + TM - 1) / TM, 320, SM_TOTAL, st>>>(a);
  CBG_LAUNCHED(CBG_K_NODE_GEMM, st);
  return 0;
}
