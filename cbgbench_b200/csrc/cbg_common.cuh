// Shared device helpers for the cbg_b200 kernels (sm_100a).
#pragma once
#include <stdlib.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include "cbg_layout.h"

#define CBG_FULL 0xffffffffu

// error plumbing (api.cu owns the storage)
void cbg_set_error(const char* fmt, ...);
#define CBG_CUDA_OK(expr)                                                        \
  do {                                                                           \
    cudaError_t _e = (expr);                                                     \
    if (_e != cudaSuccess) {                                                     \
      cbg_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #expr, cudaGetErrorString(_e)); \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

// One-time per-DEVICE setup guards: kernel attributes (opt-in shared memory) belong to a device's context, so a
// process that drives several GPUs must set them once on each.
constexpr int CBG_MAX_DEVICES = 64;
inline bool& cbg_dev_flag(bool (&flags)[CBG_MAX_DEVICES]) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= CBG_MAX_DEVICES) dev = 0;
  return flags[dev];
}

extern long long g_cbg_launches;
extern int g_cbg_prof_on;
// kernel families for the optional per-kernel CUDA-event profile (bench.py roofline)
enum CbgKernelFamily {
  CBG_K_KNN = 0, CBG_K_GATE, CBG_K_NODE_GEMM, CBG_K_X2H_K, CBG_K_X2H_V, CBG_K_H2X, CBG_K_CLASSIFIER,
  CBG_K_STEP_INIT, CBG_K_REVERSE, CBG_K_MISC, CBG_K_COUNT
};
void cbg_prof_mark(int family, int is_end, cudaStream_t st);
// bracket a kernel launch: PROF_BEGIN before it, LAUNCHED after it (surfaces launch errors,
// counts the launch, closes the profile bracket)
#define CBG_PROF_BEGIN(family, st) \
  do { if (g_cbg_prof_on) cbg_prof_mark((family), 0, (st)); } while (0)
#define CBG_LAUNCHED(family, st)                              \
  do {                                                        \
    CBG_CUDA_OK(cudaGetLastError());                          \
    g_cbg_launches += 1;                                      \
    if (g_cbg_prof_on) cbg_prof_mark((family), 1, (st));      \
  } while (0)

// Launch with (or without) programmatic stream serialization: the kernel may be scheduled before the previous kernel of
// the stream has finished; it must execute griddepcontrol.wait (cbg_tc.cuh: pdl_wait) before touching anything that
// kernel produces or still reads.  Off unless CBG_PDL=1 (measured at c2 / c1 under graph replay: no gain).
inline bool cbg_pdl_enabled() {
  static int on = -1;
  if (on < 0) { const char* e = getenv("CBG_PDL"); on = (e && e[0] == '1') ? 1 : 0; }      // opt-in: measured neutral under graph replay
  return on != 0;
}
template <typename... KArgs, typename... Args>
inline cudaError_t cbg_launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = cbg_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(CBG_FULL, v, m);
  return v;
}

// Reduce NV per-lane values across the 32 lanes of a warp with a halving butterfly.
// On return lane l holds in v[0 .. NV/32) the full (all-lane) sums of the original
// indices l*(NV/32) + i.  NV must be a power of two >= 32; 2*NV-64 shuffles... (NV-NV/32).
template <int NV>
__device__ __forceinline__ void warp_transpose_reduce(float (&v)[NV], int lane) {
#pragma unroll
  for (int s = 0; s < 5; ++s) {
    const int mask = 16 >> s;
    const int n = NV >> (s + 1);
    const bool upper = (lane & mask) != 0;
#pragma unroll
    for (int i = 0; i < n; ++i) {
      const float send = upper ? v[i] : v[i + n];
      const float keep = upper ? v[i + n] : v[i];
      v[i] = keep + __shfl_xor_sync(CBG_FULL, send, mask);
    }
  }
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// Blackwell packed fp32: FFMA2 / FMUL2 / FADD2 execute two IEEE-rn operations per issued instruction
// (sm_100a; a scalar second operand is broadcast by the hardware).  Bit-identical to the scalar forms,
// half the issue slots - the edge kernels are issue-limited, not FMA-pipe-limited.
__device__ __forceinline__ void fma4(float4& acc, const float4 w, const float s) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __ffma2_rn(make_float2(w.x, w.y), ss, make_float2(acc.x, acc.y));
  const float2 hi = __ffma2_rn(make_float2(w.z, w.w), ss, make_float2(acc.z, acc.w));
  acc = make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 add4(const float4 a, const float4 b) {
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  const float2 hi = __fadd2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w));
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
__device__ __forceinline__ float4 add4s(const float4 a, const float s) {
  const float2 ss = make_float2(s, s);
  const float2 lo = __fadd2_rn(make_float2(a.x, a.y), ss);
  const float2 hi = __fadd2_rn(make_float2(a.z, a.w), ss);
  return make_float4(lo.x, lo.y, hi.x, hi.y);
}
// a.x*b.x + a.y*b.y + a.z*b.z + a.w*b.w as (x,z | y,w) packed partial sums
__device__ __forceinline__ float dot4(const float4 a, const float4 b) {
  float2 t = __fmul2_rn(make_float2(a.x, a.y), make_float2(b.x, b.y));
  t = __ffma2_rn(make_float2(a.z, a.w), make_float2(b.z, b.w), t);
  return t.x + t.y;
}
// relu((a * rstd) * gamma + beta)
__device__ __forceinline__ float4 ln_relu4(const float4 a, const float rstd, const float4 gamma, const float4 beta) {
  const float2 rr = make_float2(rstd, rstd);
  float2 lo = __fmul2_rn(make_float2(a.x, a.y), rr), hi = __fmul2_rn(make_float2(a.z, a.w), rr);
  lo = __ffma2_rn(lo, make_float2(gamma.x, gamma.y), make_float2(beta.x, beta.y));
  hi = __ffma2_rn(hi, make_float2(gamma.z, gamma.w), make_float2(beta.z, beta.w));
  return make_float4(fmaxf(lo.x, 0.f), fmaxf(lo.y, 0.f), fmaxf(hi.x, 0.f), fmaxf(hi.y, 0.f));
}

// cooperative contiguous copy global -> shared, n floats (multiple of 4), both 16B aligned
__device__ __forceinline__ void block_copy_f4(float* dst, const float* __restrict__ src, int n_floats) {
  const int n4 = n_floats >> 2;
  for (int i = threadIdx.x; i < n4; i += blockDim.x) st4(dst + 4 * i, ldg4(src + 4 * i));
}

// node flags are stored as a float in x4.w: 0/1 = ligand bit, +2 = generate bit
__device__ __forceinline__ int node_flags(const float4 x) { return (int)x.w; }
