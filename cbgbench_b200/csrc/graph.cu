// Neighbour-list construction and the global edge gate.
//
// Replaces, on the hot path:
//   [3P] torch_geometric.nn.knn_graph   (call site /root/reference repo/modules/e3nn/unitransformer.py:79-80)
//   UniTransformer edge gate e_w         (unitransformer.py:109-112, embs/dist_emb.py:6-14,
//                                         common.py:114-133 GaussianSmearing, :151-171 MLP)
#include "cbg_kernels.cuh"

namespace {

typedef unsigned long long u64;
constexpr u64 kInfKey = ~0ull;

__device__ __forceinline__ u64 bitonic_sort32(u64 v, int lane) {
#pragma unroll
  for (int k = 2; k <= 32; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      const u64 o = __shfl_xor_sync(CBG_FULL, v, j);
      const bool up = (lane & k) == 0;     // k == 32 -> always ascending
      const bool lower = (lane & j) == 0;
      const u64 mn = v < o ? v : o, mx = v < o ? o : v;
      v = (lower == up) ? mn : mx;
    }
  }
  return v;
}

__device__ __forceinline__ u64 bitonic_merge32(u64 v, int lane) {
#pragma unroll
  for (int j = 16; j > 0; j >>= 1) {
    const u64 o = __shfl_xor_sync(CBG_FULL, v, j);
    const u64 mn = v < o ? v : o, mx = v < o ? o : v;
    v = ((lane & j) == 0) ? mn : mx;
  }
  return v;
}

// One CTA = one (graph, chunk of 64 centres).  The whole graph's coordinates sit in shared
// memory; a warp owns one centre at a time and keeps the running 32 best (key = d2 bits:idx)
// sorted across its lanes; candidate batches of 32 are bitonic-sorted and merged.
// Squared distance uses individually rounded mul/add (no FMA) so the selection is
// bit-identical to the oracle (oracle/graph_ops.py).
__global__ void __launch_bounds__(256) knn_kernel(const float4* __restrict__ x4,
                                                  const int* __restrict__ graph_ptr, int k, int mode,
                                                  float r2max, int static_only, const int* __restrict__ snbr,
                                                  int* __restrict__ nbr) {
  extern __shared__ float4 xs[];
  const int g = blockIdx.x;          // graphs on x (2^31 - 1 blocks), 64-centre chunks on y (<= 200)
  const int s = graph_ptr[g];
  const int n = graph_ptr[g + 1] - s;
  const int c0 = blockIdx.y * 64;
  if (c0 >= n) return;
  for (int i = threadIdx.x; i < n; i += blockDim.x) xs[i] = x4[s + i];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c1 = min(n, c0 + 64);
  for (int c = c0 + warp; c < c1; c += 8) {
    const float4 xc = xs[c];
    u64 best = kInfKey;
    if (static_only && (node_flags(xc) & 2)) {          // moving centre: no static edges
      nbr[(size_t)(s + c) * CBG_KMAX + lane] = -1;
      continue;
    }
    // Incremental search for a non-moving centre: its 32 nearest NON-moving neighbours are known (static list,
    // same keys, already sorted), so only the moving atoms of the graph have to be merged in.
    const bool incremental = (snbr != nullptr) && !(node_flags(xc) & 2);
    if (incremental) {
      const int js = snbr[(size_t)(s + c) * CBG_KMAX + lane];
      if (js >= 0) {
        const float4 xj = xs[js - s];
        const float dx = xc.x - xj.x, dy = xc.y - xj.y, dz = xc.z - xj.z;
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        best = ((u64)__float_as_uint(d2) << 32) | (u64)(unsigned)(js - s);
      }
    }
    for (int base = 0; base < n; base += 32) {
      const int j = base + lane;
      u64 key = kInfKey;
      if (j < n && j != c && !(static_only && (node_flags(xs[j]) & 2)) &&
          !(incremental && !(node_flags(xs[j]) & 2))) {
        const float4 xj = xs[j];
        const float dx = xc.x - xj.x, dy = xc.y - xj.y, dz = xc.z - xj.z;
        const float d2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
        key = ((u64)__float_as_uint(d2) << 32) | (u64)(unsigned)j;
      }
      const u64 worst = __shfl_sync(CBG_FULL, best, 31);
      if (!__any_sync(CBG_FULL, key < worst)) continue;
      key = bitonic_sort32(key, lane);
      const u64 rev = __shfl_sync(CBG_FULL, key, 31 - lane);
      best = best < rev ? best : rev;
      best = bitonic_merge32(best, lane);
    }
    int out = -1;
    if (best != kInfKey && lane < k) {
      const float d2 = __uint_as_float((unsigned)(best >> 32));
      if (mode == CBG_MODE_KNN || d2 <= r2max) out = s + (int)(unsigned)(best & 0xffffffffull);
    }
    nbr[(size_t)(s + c) * CBG_KMAX + lane] = out;
  }
}

// Edge gate: one thread per (node, slot).  e_w = sigmoid(W1 . relu(LN(W0 g(d) + b0)) + b1).
constexpr int kGateSmemFloats = 20 * 160 + 160 + 320 + 160 + 32;

__device__ __forceinline__ float gate_value(const float* sm, const float4 xi, const float4 xj);

// `glist` (optional, needs ew_static): instead of computing the gates of the moving edges in place (a few
// scattered lanes per warp), append their slot indices to glist[64..] (count in glist[0]) and leave the
// arithmetic to edge_gate_list_kernel, which runs it on dense warps.
__global__ void __launch_bounds__(128) edge_gate_kernel(const float* __restrict__ gw,  // GATE_W0T..GATE_RBF
                                                        const float4* __restrict__ x4,
                                                        const int* __restrict__ nbr, long long n_slots,
                                                        const float* __restrict__ ew_static,
                                                        int* __restrict__ glist,
                                                        float* __restrict__ ew,
                                                        unsigned char* __restrict__ full_static) {
  __shared__ __align__(16) float sm[kGateSmemFloats];
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // one warp = one node's 32 slots
  const bool in_range = idx < n_slots;   // n_slots is a multiple of 32: whole warps are in or out
  int j = -1, i = 0;
  float4 xi = make_float4(0.f, 0.f, 0.f, 0.f), xj = xi;
  bool need = false;
  if (in_range) {
    j = nbr[idx];
    i = (int)(idx / CBG_KMAX);
    xi = x4[i];
    xj = x4[j >= 0 ? j : i];
    need = j >= 0;
    if (ew_static != nullptr) {
      // gate of an edge between two non-generated atoms never changes over the diffusion steps: it was
      // computed once per batch for the node's static-only neighbour list; the p-th static edge of the
      // current list is the p-th entry of that list (prefix property, see edge.cu: edge_setup)
      const bool is_static = j >= 0 && ((node_flags(xi) | node_flags(xj)) & 2) == 0;
      const unsigned sm_mask = __ballot_sync(CBG_FULL, is_static);
      if (full_static != nullptr && (threadIdx.x & 31) == 0) full_static[i] = sm_mask == 0xffffffffu;
      if (is_static) {
        ew[idx] = ew_static[(size_t)i * CBG_KMAX + __popc(sm_mask & ((1u << (threadIdx.x & 31)) - 1u))];
        need = false;
      }
    }
    if (j < 0) ew[idx] = 0.f;
  }
  if (glist != nullptr) {
    const unsigned m = __ballot_sync(CBG_FULL, need);
    if (m) {
      const int lane = threadIdx.x & 31;
      int base = 0;
      if (lane == 0) base = atomicAdd(glist, __popc(m));
      base = __shfl_sync(CBG_FULL, base, 0);
      if (need) glist[64 + base + __popc(m & ((1u << lane) - 1u))] = (int)idx;
    }
    return;
  }
  if (!__syncthreads_or(need ? 1 : 0)) return;     // nothing to compute in this CTA: skip the weight staging
  block_copy_f4(sm, gw, kGateSmemFloats);
  __syncthreads();
  if (!need) return;
  ew[idx] = gate_value(sm, xi, xj);
}

// gates of the compacted moving edges (any order: every entry owns its slot)
__global__ void __launch_bounds__(128) edge_gate_list_kernel(const float* __restrict__ gw,
                                                             const float4* __restrict__ x4,
                                                             const int* __restrict__ nbr,
                                                             const int* __restrict__ glist,
                                                             float* __restrict__ ew) {
  __shared__ __align__(16) float sm[kGateSmemFloats];
  const int n = glist[0];
  if ((long long)blockIdx.x * blockDim.x >= n) return;
  block_copy_f4(sm, gw, kGateSmemFloats);
  __syncthreads();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n) return;
  const int idx = glist[64 + t];
  ew[idx] = gate_value(sm, x4[idx / CBG_KMAX], x4[nbr[idx]]);
}

// gates of the listed rows (DiffBP CoM head: generated atoms only); 4 rows per CTA
__global__ void __launch_bounds__(128) edge_gate_rows_kernel(const float* __restrict__ gw,
                                                             const float4* __restrict__ x4,
                                                             const int* __restrict__ nbr,
                                                             const int* __restrict__ row_idx, int n_rows,
                                                             float* __restrict__ ew) {
  __shared__ __align__(16) float sm[kGateSmemFloats];
  block_copy_f4(sm, gw, kGateSmemFloats);
  __syncthreads();
  const int r = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (r >= n_rows) return;
  const int i = row_idx[r];
  const size_t idx = (size_t)i * CBG_KMAX + (threadIdx.x & 31);
  const int j = nbr[idx];
  ew[idx] = j >= 0 ? gate_value(sm, x4[i], x4[j]) : 0.f;
}

__device__ __forceinline__ float gate_value(const float* sm, const float4 xi, const float4 xj) {
  const float* w0t = sm;                 // [20][160]
  const float* b0 = sm + 3200;
  const float* gamma = b0 + 160;
  const float* beta = gamma + 160;
  const float* w1 = beta + 160;
  const float* rbf = w1 + 160;           // offsets[20], coeff, b1
  const float rx = xi.x - xj.x, ry = xi.y - xj.y, rz = xi.z - xj.z;
  const float d = sqrtf(rx * rx + ry * ry + rz * rz);
  float g[CBG_NRBF];
  const float coeff = rbf[20];
#pragma unroll
  for (int m = 0; m < CBG_NRBF; ++m) { const float u = d - rbf[m]; g[m] = expf(coeff * u * u); }
  float4 hid[40];
#pragma unroll
  for (int u = 0; u < 40; ++u) hid[u] = ld4(b0 + 4 * u);
#pragma unroll
  for (int m = 0; m < CBG_NRBF; ++m) {
#pragma unroll
    for (int u = 0; u < 40; ++u) fma4(hid[u], ld4(w0t + m * 160 + 4 * u), g[m]);
  }
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 40; ++u) s += (hid[u].x + hid[u].y) + (hid[u].z + hid[u].w);
  const float mean = s * (1.f / 160.f);
  float q = 0.f;
#pragma unroll
  for (int u = 0; u < 40; ++u) {
    hid[u].x -= mean; hid[u].y -= mean; hid[u].z -= mean; hid[u].w -= mean;
    q += (hid[u].x * hid[u].x + hid[u].y * hid[u].y) + (hid[u].z * hid[u].z + hid[u].w * hid[u].w);
  }
  const float rstd = 1.f / sqrtf(q * (1.f / 160.f) + 1e-5f);
  float o = rbf[21];
#pragma unroll
  for (int u = 0; u < 40; ++u) {
    const float4 ga = ld4(gamma + 4 * u), be = ld4(beta + 4 * u), ww = ld4(w1 + 4 * u);
    o = fmaf(ww.x, fmaxf(fmaf(hid[u].x * rstd, ga.x, be.x), 0.f), o);
    o = fmaf(ww.y, fmaxf(fmaf(hid[u].y * rstd, ga.y, be.y), 0.f), o);
    o = fmaf(ww.z, fmaxf(fmaf(hid[u].z * rstd, ga.z, be.z), 0.f), o);
    o = fmaf(ww.w, fmaxf(fmaf(hid[u].w * rstd, ga.w, be.w), 0.f), o);
  }
  return 1.f / (1.f + expf(-o));
}

// ---- receptive-field depth ------------------------------------------------------------------------
// X2H(l) must produce h for D_l, with D_{L-1} = seeds U nbr(seeds) U cls (H2X of every layer reads the
// new h of the generated atoms and of their neighbours; the classifier reads the last h of cls) and
// D_{l-1} = D_l U nbr(D_l).  depth[i] = max l with i in D_l (down to -1 for the Pj planes of layer 0).
__global__ void depth_seed_kernel(const int* __restrict__ nbr, const int* __restrict__ seed_idx, int n_seed,
                                  const int* __restrict__ cls_idx, int n_cls, int top, int* __restrict__ depth) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n_seed * (CBG_KMAX + 1)) {
    const int g = seed_idx[t / (CBG_KMAX + 1)], s = t % (CBG_KMAX + 1);
    const int j = (s == CBG_KMAX) ? g : nbr[(size_t)g * CBG_KMAX + s];
    if (j >= 0) depth[j] = top;      // benign race: every writer stores the same value
  }
  if (t < n_cls) depth[cls_idx[t]] = top;
}

// one CTA per graph: relax level by level inside the graph's node range
__global__ void __launch_bounds__(1024) depth_relax_kernel(const int* __restrict__ nbr, const int* __restrict__ graph_ptr,
                                                          int top, int* __restrict__ depth) {
  const int s = graph_ptr[blockIdx.x], e = graph_ptr[blockIdx.x + 1];
  for (int l = top; l >= 0; --l) {
    for (int t = threadIdx.x; t < (e - s) * CBG_KMAX; t += blockDim.x) {
      const int i = s + t / CBG_KMAX;
      if (depth[i] >= l) {
        const int j = nbr[(size_t)i * CBG_KMAX + (t % CBG_KMAX)];
        if (j >= 0 && depth[j] < l - 1) atomicMax(&depth[j], l - 1);
      }
    }
    __syncthreads();
  }
}

// counting sort of the nodes by decreasing depth (bins top .. -1; depth -2 nodes are dropped)
__global__ void depth_hist_kernel(const int* __restrict__ depth, long long n, int top, int* __restrict__ hist) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && depth[i] >= -1) atomicAdd(&hist[top - depth[i]], 1);
}
__global__ void depth_scan_kernel(int top, int* __restrict__ hist, int* __restrict__ cursor, int* __restrict__ cnt_ge) {
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int b = 0; b <= top + 1; ++b) {          // bin b holds depth top - b
      cursor[b] = acc;
      acc += hist[b];
      cnt_ge[(top - b) + 1] = acc;                // nodes with depth >= top - b
    }
  }
}
__global__ void depth_scatter_kernel(const int* __restrict__ depth, long long n, int top, int* __restrict__ cursor,
                                     int* __restrict__ order) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && depth[i] >= -1) order[atomicAdd(&cursor[top - depth[i]], 1)] = (int)i;
}

}  // namespace

int cbg_launch_depth(const int* nbr, const int* graph_ptr, int n_graphs, int max_graph_nodes, long long n_nodes,
                     const int* seed_idx, int n_seed, const int* cls_idx, int n_cls, int num_layers,
                     int* depth, int* order, int* cnt_ge, cudaStream_t st) {
  if (n_nodes <= 0 || num_layers <= 0) return 0;
  if (num_layers > 30) { cbg_set_error("too many layers for receptive-field pruning"); return 1; }
  const int top = num_layers - 1;
  int* hist = cnt_ge + 32;      // scratch behind the counts: hist[32], cursor[32]
  int* cursor = cnt_ge + 64;
  CBG_CUDA_OK(cudaMemsetAsync(depth, 0xFE, (size_t)n_nodes * sizeof(int), st));   // 0xFEFEFEFE = -16843010 < -1
  CBG_CUDA_OK(cudaMemsetAsync(cnt_ge, 0, 96 * sizeof(int), st));
  const int work = (n_seed * (CBG_KMAX + 1) > n_cls) ? n_seed * (CBG_KMAX + 1) : n_cls;
  if (work > 0) {
    CBG_PROF_BEGIN(CBG_K_MISC, st);
    depth_seed_kernel<<<(work + 255) / 256, 256, 0, st>>>(nbr, seed_idx, n_seed, cls_idx, n_cls, top, depth);
    CBG_LAUNCHED(CBG_K_MISC, st);
  }
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  depth_relax_kernel<<<n_graphs, 1024, 0, st>>>(nbr, graph_ptr, top, depth);
  CBG_LAUNCHED(CBG_K_MISC, st);
  const unsigned nb = (unsigned)((n_nodes + 255) / 256);
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  depth_hist_kernel<<<nb, 256, 0, st>>>(depth, n_nodes, top, hist);
  CBG_LAUNCHED(CBG_K_MISC, st);
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  depth_scan_kernel<<<1, 32, 0, st>>>(top, hist, cursor, cnt_ge);
  CBG_LAUNCHED(CBG_K_MISC, st);
  CBG_PROF_BEGIN(CBG_K_MISC, st);
  depth_scatter_kernel<<<nb, 256, 0, st>>>(depth, n_nodes, top, cursor, order);
  CBG_LAUNCHED(CBG_K_MISC, st);
  (void)max_graph_nodes;
  return 0;
}

int cbg_launch_knn(const float4* x4, const int* graph_ptr, int n_graphs, int max_graph_nodes, int mode,
                   int k, float r_max, int static_only, const int* snbr, int* nbr, cudaStream_t st) {
  if (n_graphs <= 0) return 0;
  if (k < 1 || k > CBG_KMAX) { cbg_set_error("k=%d outside [1,%d]", k, CBG_KMAX); return 1; }
  const size_t smem = (size_t)max_graph_nodes * sizeof(float4);
  if (smem > 200 * 1024) {
    cbg_set_error("graph with %d atoms exceeds the %d-atom shared-memory limit of the neighbour search",
                  max_graph_nodes, 200 * 1024 / 16);
    return 1;
  }
  static size_t smem_attr_dev[CBG_MAX_DEVICES] = {};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= CBG_MAX_DEVICES) dev = 0;
  size_t& smem_attr = smem_attr_dev[dev];
  if (smem > 48 * 1024 && smem > smem_attr) {
    CBG_CUDA_OK(cudaFuncSetAttribute(knn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_attr = smem;
  }
  dim3 grid(n_graphs, (max_graph_nodes + 63) / 64);
  CBG_PROF_BEGIN(CBG_K_KNN, st);
  knn_kernel<<<grid, 256, smem, st>>>(x4, graph_ptr, k, mode, r_max * r_max, static_only, snbr, nbr);
  CBG_LAUNCHED(CBG_K_KNN, st);
  if (cudaError_t e = cudaPeekAtLastError()) { cbg_set_error("knn_kernel launch failed: %s", cudaGetErrorString(e)); return 2; }
  return 0;
}

int cbg_launch_edge_gate_rows(const float* blob_global, const float4* x4, const int* nbr, const int* row_idx,
                              int n_rows, float* ew, cudaStream_t st) {
  if (n_rows <= 0) return 0;
  const float* gw = blob_global + cbg_layout::global_offset(CBG_GF_GATE_W0T);
  CBG_PROF_BEGIN(CBG_K_GATE, st);
  edge_gate_rows_kernel<<<(n_rows + 3) / 4, 128, 0, st>>>(gw, x4, nbr, row_idx, n_rows, ew);
  CBG_LAUNCHED(CBG_K_GATE, st);
  return 0;
}

int cbg_launch_edge_gate(const float* blob_global, const float4* x4, const int* nbr, long long n_nodes,
                         const float* ew_static, int* glist, float* ew, cudaStream_t st, unsigned char* full_static) {
  const long long n_slots = n_nodes * CBG_KMAX;
  if (n_slots == 0) return 0;
  const float* gw = blob_global + cbg_layout::global_offset(CBG_GF_GATE_W0T);
  CBG_PROF_BEGIN(CBG_K_GATE, st);
  const unsigned grid = (unsigned)((n_slots + 127) / 128);
  if (ew_static == nullptr) glist = nullptr;
  if (glist) CBG_CUDA_OK(cudaMemsetAsync(glist, 0, sizeof(int), st));
  edge_gate_kernel<<<grid, 128, 0, st>>>(gw, x4, nbr, n_slots, ew_static, glist, ew, ew_static ? full_static : nullptr);
  if (glist) {
    edge_gate_list_kernel<<<grid, 128, 0, st>>>(gw, x4, nbr, glist, ew);
    g_cbg_launches += 1;
  }
  CBG_LAUNCHED(CBG_K_GATE, st);
  return 0;
}
