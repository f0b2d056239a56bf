// Device-side batch construction for sampling.
//
// The reference evaluates `dataset[i]` num_samples times per pocket (sample.py:177), i.e. 100-200 passes of the
// Python transform list of configs/*/test/*.yml (featurize_protein_fa, center_pos, assign_molsize / assign_gensize,
// assign_atomtype / assign_genatomtype, assign_molpos / assign_genpos, merge) followed by PyG's collate.  Here the
// same batch is produced by three kernels from the raw pocket arrays; random numbers stay with the caller
// (include/cbg_b200.h), so identical draws give identical batches.
//
//   pocket_stats_kernel   centre of the pocket (translation.py:11-24,36-50) and its "space size"
//                         = median of the 10 largest pairwise atom distances (init_lig.py:247-250)
//   ligand_sizes_kernel   size prior: bin by space size (init_lig.py:47-52), numpy's legacy choice(values, p) with the
//                         caller's uniform draw (init_lig.py:27-31), assign_gensize's context rule (:269-271), offsets
//   build_batch_kernel    one CTA per sample: centred + featurised protein copy (protein_featurizer.py:19-30), ligand
//                         rows = [context atoms | generated atoms] with uniform / absorbing types (init_lig.py:22-26,
//                         299-341, 373-401) and Gaussian / zero-mean Gaussian positions (:404-457), flags, graph ids
#include <math.h>
#include "cbg_kernels.cuh"

namespace {

constexpr int kTop = 10;
constexpr int kStatThreads = 256;

__device__ __forceinline__ double block_sum_double(double v, double* scratch /*[32]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int m = 16; m >= 1; m >>= 1) v += __shfl_xor_sync(CBG_FULL, v, m);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  double t = 0.0;
  for (int w = 0; w < nwarp; ++w) t += scratch[w];      // fixed order: deterministic
  return t;
}

// centre_mode 0: mean of the pocket atoms (center_pos(protein), center_whole_pos without ligand) - the space size is
//                then measured on the centred coordinates, as assign_molsize sees them;
// centre_mode 1: mean of the pocket's context ligand atoms, 0 if it has none (center_pos(ligand, mask ctx_flag)) -
//                assign_gensize runs before the centring, so the space size uses the raw coordinates.
__global__ void __launch_bounds__(kStatThreads) pocket_stats_kernel(const float* __restrict__ prot_pos,
                                                                    const int* __restrict__ prot_ptr,
                                                                    const float* __restrict__ ctx_pos,
                                                                    const int* __restrict__ ctx_ptr, int centre_mode,
                                                                    float* __restrict__ space_size,
                                                                    float* __restrict__ centre) {
  __shared__ double s_red[32];
  __shared__ float s_top[kStatThreads][kTop];
  __shared__ unsigned long long s_best[kStatThreads / 32];
  __shared__ float s_c[3];
  const int p = blockIdx.x, tid = threadIdx.x;
  const int p0 = prot_ptr[p], P = prot_ptr[p + 1] - p0;
  const float* x = prot_pos + (size_t)p0 * 3;
  {
    const float* src = x;
    int n = P;
    if (centre_mode == 1) { const int c0 = ctx_ptr[p]; n = ctx_ptr[p + 1] - c0; src = ctx_pos + (size_t)c0 * 3; }
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int a = tid; a < n; a += blockDim.x) { sx += src[3 * a]; sy += src[3 * a + 1]; sz += src[3 * a + 2]; }
    sx = block_sum_double(sx, s_red); sy = block_sum_double(sy, s_red); sz = block_sum_double(sz, s_red);
    if (tid == 0) {
      s_c[0] = n > 0 ? (float)(sx / n) : 0.f; s_c[1] = n > 0 ? (float)(sy / n) : 0.f; s_c[2] = n > 0 ? (float)(sz / n) : 0.f;
      centre[3 * p] = s_c[0]; centre[3 * p + 1] = s_c[1]; centre[3 * p + 2] = s_c[2];
    }
    __syncthreads();
  }
  const float cx = centre_mode == 0 ? s_c[0] : 0.f, cy = centre_mode == 0 ? s_c[1] : 0.f, cz = centre_mode == 0 ? s_c[2] : 0.f;
  // per-thread top-10 of the squared pair distances (non-fused arithmetic, like a scalar CPU loop)
  float top[kTop];
#pragma unroll
  for (int k = 0; k < kTop; ++k) top[k] = -1.f;
  for (int i = tid; i < P; i += blockDim.x) {
    const float xi = __fsub_rn(x[3 * i], cx), yi = __fsub_rn(x[3 * i + 1], cy), zi = __fsub_rn(x[3 * i + 2], cz);
    for (int j = i + 1; j < P; ++j) {
      const float dx = __fsub_rn(xi, __fsub_rn(x[3 * j], cx)), dy = __fsub_rn(yi, __fsub_rn(x[3 * j + 1], cy)),
                  dz = __fsub_rn(zi, __fsub_rn(x[3 * j + 2], cz));
      const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
      if (d2 > top[kTop - 1]) {
        top[kTop - 1] = d2;
#pragma unroll
        for (int k = kTop - 1; k > 0; --k)
          if (top[k] > top[k - 1]) { const float t = top[k]; top[k] = top[k - 1]; top[k - 1] = t; }
      }
    }
  }
#pragma unroll
  for (int k = 0; k < kTop; ++k) s_top[tid][k] = top[k];
  __syncthreads();
  // k-way merge by repeated block arg-max: the (target+1)-th largest pair distance is the lower median of the top 10
  const long long n_pairs = (long long)P * (P - 1) / 2;
  const int m = n_pairs < kTop ? (int)n_pairs : kTop;
  if (m <= 0) { if (tid == 0) space_size[p] = nanf(""); return; }     // torch.median of an empty tensor has no value
  const int target = m - 1 - (m - 1) / 2;
  int head = 0;
  float answer = 0.f;
  for (int r = 0; r <= target; ++r) {
    const float v = head < kTop ? s_top[tid][head] : -1.f;
    unsigned long long key = v >= 0.f ? (((unsigned long long)__float_as_uint(v) << 32) | (unsigned)(0xffffu - tid)) : 0ull;
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) { const unsigned long long o = __shfl_xor_sync(CBG_FULL, key, s); key = o > key ? o : key; }
    if ((tid & 31) == 0) s_best[tid >> 5] = key;
    __syncthreads();
    unsigned long long best = 0ull;
    for (int w = 0; w < kStatThreads / 32; ++w) best = s_best[w] > best ? s_best[w] : best;
    __syncthreads();
    if ((int)(0xffffu - (unsigned)(best & 0xffffu)) == tid) ++head;
    answer = __uint_as_float((unsigned)(best >> 32));
  }
  if (tid == 0) space_size[p] = __fsqrt_rn(answer);
}

struct SizeArgs {
  const double* bounds; int n_bounds;
  const int* bin_ptr; const int* values; const double* cdf;
  const float* space_size; int n_pockets; int repeat;
  const double* u; const int* ctx_ptr; const int* extra;
  int* n_lig; int* lig_ptr;
};

__global__ void __launch_bounds__(256) ligand_sizes_kernel(SizeArgs a) {
  const int S = a.n_pockets * a.repeat;
  for (int s = threadIdx.x; s < S; s += blockDim.x) {
    const int p = s / a.repeat;
    const double size = (double)a.space_size[p];
    int bin = a.n_bounds;
    for (int i = 0; i < a.n_bounds; ++i)
      if (a.bounds[i] > size) { bin = i; break; }
    const int lo = a.bin_ptr[bin], hi = a.bin_ptr[bin + 1];
    const double u = a.u[s];
    int l = lo, h = hi;                       // searchsorted(cdf, u, side='right'): first index with cdf > u
    while (l < h) { const int mid = (l + h) >> 1; if (a.cdf[mid] <= u) l = mid + 1; else h = mid; }
    if (l >= hi) l = hi - 1;
    int n = a.values[l];
    if (a.ctx_ptr) {
      const int c = a.ctx_ptr[p + 1] - a.ctx_ptr[p];
      if (n <= c) n = c + a.extra[s];         // init_lig.py:269-271
    }
    a.n_lig[s] = n;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int s = 0; s < S; ++s) { a.lig_ptr[s] = acc; acc += a.n_lig[s]; }
    a.lig_ptr[S] = acc;
  }
}

struct BuildArgs {
  const float* prot_pos; const int* prot_element; const unsigned char* prot_backbone; const int* prot_aa;
  const int* prot_ptr; int n_pockets; int repeat;
  const float* centre;
  const float* ctx_pos; const int* ctx_type; const int* ctx_ptr;
  const int* lig_ptr;
  const float* pos_noise; const float* type_u;
  int num_classes, type_dist, pos_dist;
  float* o_prot_pos; float* o_prot_feat; long long* o_prot_aa; long long* o_prot_batch; float* o_prot_tr;
  float* o_lig_pos; long long* o_lig_type; long long* o_lig_batch; unsigned char* o_lig_ctx; unsigned char* o_lig_gen;
};

__global__ void __launch_bounds__(256) build_batch_kernel(BuildArgs a) {
  __shared__ double s_red[32];
  const int s = blockIdx.x, tid = threadIdx.x;
  const int p = s / a.repeat;
  const int p0 = a.prot_ptr[p], P = a.prot_ptr[p + 1] - p0;
  const float cx = a.centre[3 * p], cy = a.centre[3 * p + 1], cz = a.centre[3 * p + 2];
  const size_t po = (size_t)a.repeat * p0 + (size_t)(s - p * a.repeat) * P;      // every sample of a pocket has P atoms
  for (int i = tid; i < P; i += blockDim.x) {
    const size_t o = po + i;
    const float* x = a.prot_pos + (size_t)(p0 + i) * 3;
    a.o_prot_pos[3 * o] = __fsub_rn(x[0], cx); a.o_prot_pos[3 * o + 1] = __fsub_rn(x[1], cy); a.o_prot_pos[3 * o + 2] = __fsub_rn(x[2], cz);
    a.o_prot_tr[3 * o] = cx; a.o_prot_tr[3 * o + 1] = cy; a.o_prot_tr[3 * o + 2] = cz;
    const int el = a.prot_element[p0 + i];
    float* f = a.o_prot_feat + 7 * o;          // H, C, N, O, S, Se one-hot + backbone flag (repo/utils/protein/constants.py)
    f[0] = el == 1; f[1] = el == 6; f[2] = el == 7; f[3] = el == 8; f[4] = el == 16; f[5] = el == 34;
    f[6] = a.prot_backbone[p0 + i] ? 1.f : 0.f;
    a.o_prot_aa[o] = a.prot_aa[p0 + i];
    a.o_prot_batch[o] = s;
  }
  const int l0 = a.lig_ptr[s], L = a.lig_ptr[s + 1] - l0;
  const int c0 = a.ctx_ptr ? a.ctx_ptr[p] : 0, C = a.ctx_ptr ? a.ctx_ptr[p + 1] - c0 : 0;
  float mx = 0.f, my = 0.f, mz = 0.f;
  if (a.pos_dist == 1) {                       // zero_mean_gaussian: subtract the sample's mean noise (init_lig.py:415-417)
    double sx = 0.0, sy = 0.0, sz = 0.0;
    for (int l = tid; l < L; l += blockDim.x) {
      const float* nz = a.pos_noise + (size_t)(l0 + l) * 3;
      sx += nz[0]; sy += nz[1]; sz += nz[2];
    }
    sx = block_sum_double(sx, s_red); sy = block_sum_double(sy, s_red); sz = block_sum_double(sz, s_red);
    if (L > 0) { mx = (float)(sx / L); my = (float)(sy / L); mz = (float)(sz / L); }
  }
  for (int l = tid; l < L; l += blockDim.x) {
    const size_t o = (size_t)l0 + l;
    const bool is_ctx = l < C;
    float x, y, z;
    long long type;
    if (is_ctx) {
      const float* cp = a.ctx_pos + (size_t)(c0 + l) * 3;
      x = __fsub_rn(cp[0], cx); y = __fsub_rn(cp[1], cy); z = __fsub_rn(cp[2], cz);
      type = a.ctx_type[c0 + l];
    } else {
      const float* nz = a.pos_noise + o * 3;
      x = __fsub_rn(nz[0], mx); y = __fsub_rn(nz[1], my); z = __fsub_rn(nz[2], mz);
      type = 0;                                // absorbing state / placeholder of the 'zeros' distribution
      if (a.type_dist == 0) {                  // Gumbel arg-max over zero logits (init_lig.py:22-26), first maximum wins
        const float* u = a.type_u + o * a.num_classes;
        float best = -INFINITY;
        for (int k = 0; k < a.num_classes; ++k) {
          const float gk = -logf(__fadd_rn(-logf(__fadd_rn(u[k], 1e-30f)), 1e-30f));
          if (gk > best) { best = gk; type = k; }
        }
      }
    }
    a.o_lig_pos[3 * o] = x; a.o_lig_pos[3 * o + 1] = y; a.o_lig_pos[3 * o + 2] = z;
    a.o_lig_type[o] = type;
    a.o_lig_batch[o] = s;
    if (a.o_lig_ctx) a.o_lig_ctx[o] = is_ctx;
    if (a.o_lig_gen) a.o_lig_gen[o] = !is_ctx;
  }
}

}  // namespace

int cbg_launch_pocket_stats(const float* prot_pos, const int* prot_ptr, int n_pockets, const float* ctx_pos,
                            const int* ctx_ptr, int centre_mode, float* space_size, float* centre, cudaStream_t st) {
  if (n_pockets <= 0) return 0;
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  pocket_stats_kernel<<<n_pockets, kStatThreads, 0, st>>>(prot_pos, prot_ptr, ctx_pos, ctx_ptr, centre_mode, space_size, centre);
  CBG_LAUNCHED(CBG_K_MISC, st);
  return 0;
}

int cbg_launch_ligand_sizes(const double* bounds, int n_bounds, const int* bin_ptr, const int* values, const double* cdf,
                            const float* space_size, int n_pockets, int repeat, const double* u, const int* ctx_ptr,
                            const int* extra, int* n_lig, int* lig_ptr, cudaStream_t st) {
  if (n_pockets <= 0 || repeat <= 0) return 0;
  SizeArgs a{bounds, n_bounds, bin_ptr, values, cdf, space_size, n_pockets, repeat, u, ctx_ptr, extra, n_lig, lig_ptr};
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  ligand_sizes_kernel<<<1, 256, 0, st>>>(a);
  CBG_LAUNCHED(CBG_K_MISC, st);
  return 0;
}

int cbg_launch_build_batch(const float* prot_pos, const int* prot_element, const unsigned char* prot_backbone,
                           const int* prot_aa, const int* prot_ptr, int n_pockets, int repeat, const float* centre,
                           const float* ctx_pos, const int* ctx_type, const int* ctx_ptr, const int* lig_ptr,
                           const float* pos_noise, const float* type_u, int num_classes, int type_dist, int pos_dist,
                           float* o_prot_pos, float* o_prot_feat, long long* o_prot_aa, long long* o_prot_batch,
                           float* o_prot_tr, float* o_lig_pos, long long* o_lig_type, long long* o_lig_batch,
                           unsigned char* o_lig_ctx, unsigned char* o_lig_gen, cudaStream_t st) {
  if (n_pockets <= 0 || repeat <= 0) return 0;
  BuildArgs a{prot_pos, prot_element, prot_backbone, prot_aa, prot_ptr, n_pockets, repeat, centre, ctx_pos, ctx_type,
              ctx_ptr, lig_ptr, pos_noise, type_u, num_classes, type_dist, pos_dist, o_prot_pos, o_prot_feat, o_prot_aa,
              o_prot_batch, o_prot_tr, o_lig_pos, o_lig_type, o_lig_batch, o_lig_ctx, o_lig_gen};
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  build_batch_kernel<<<n_pockets * repeat, 256, 0, st>>>(a);
  CBG_LAUNCHED(CBG_K_MISC, st);
  return 0;
}
