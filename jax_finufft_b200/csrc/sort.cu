// sort.cu -- bin-sort of the non-uniform points (setpts for types 1/2, and both point sets of
// type 3).  Replaces calc_bin_size_noghost_* + thrust scans + calc_inverse_of_global_sort_index_*
// + calc_subprob_* + map_b_into_subprob_* (V/src/cuda/3d/spreadinterp3d.cuh:28-84,
// V/src/cuda/precision_independent.cu:37-91, V/src/cuda/3d/spread3d_wrapper.cu:418-506).
//
// Contract kept from the reference: bin id = binx + biny*nbinx + binz*nbinx*nbiny
// with bin_d = clamp(floor(fold_rescale(x_d)/binsize_d)); bins concatenated by exclusive scan;
// order inside a bin unspecified.  B200-specific choices: warp-aggregated histogram atomics
// (one atomic per distinct bin per warp -- clustered inputs no longer serialise on one counter),
// a device-side scan and subproblem list with NO host readback / stream sync, and the folded
// coordinates written out in sorted order (12 B/pt once) so the spread/interp kernels stream
// them coalesced instead of gathering x[idx[i]] (three 32-B sectors per point) every transform.
//
// HBM bytes per point (float, 3-D): pass 1 reads 12, writes 8; pass 2 reads 12+8, writes 16
// => 56 B/pt algorithmic.
#include "plan.h"

namespace b2n {

int dev_alloc(void **p, size_t bytes, cudaStream_t st) {
  if (bytes == 0) bytes = 16;
  cudaError_t e = cudaMallocAsync(p, bytes, st);
  if (e != cudaSuccess) {
    fprintf(stderr, "[b200nufft] cudaMallocAsync(%zu) failed: %s\n", bytes, cudaGetErrorString(e));
    *p = nullptr;
    cudaGetLastError();
    return B2N_ERR_ALLOC;
  }
  return 0;
}
void dev_free(void *p, cudaStream_t st) {
  if (p) cudaFreeAsync(p, st);
}

// ------------------------------------------------------------------------------- scan (int32)
constexpr int SCAN_T = 256;
constexpr int SCAN_E = 8;  // elements per thread
constexpr int SCAN_CH = SCAN_T * SCAN_E;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total, int *sm /*[32]*/) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) sm[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int s = lane < (blockDim.x >> 5) ? sm[lane] : 0;
    int t = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      int n = __shfl_up_sync(0xffffffffu, t, o);
      if (lane >= o) t += n;
    }
    sm[lane] = t - s;  // exclusive warp offsets
    if (lane == 31) sm[32] = t;
  }
  __syncthreads();
  int res = sm[wid] + inc - v;
  *total = sm[32];
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_blocks(const int32_t *__restrict__ in,
                                                          int32_t *__restrict__ out, int64_t n,
                                                          int32_t *__restrict__ bsum) {
  __shared__ int sm[33];
  const int64_t base = (int64_t)blockIdx.x * SCAN_CH + (int64_t)threadIdx.x * SCAN_E;
  int v[SCAN_E];
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    v[k] = (base + k < n) ? in[base + k] : 0;
    s += v[k];
  }
  int total;
  int off = block_exclusive_scan(s, &total, sm);
#pragma unroll
  for (int k = 0; k < SCAN_E; k++) {
    if (base + k < n) out[base + k] = off;
    off += v[k];
  }
  if (threadIdx.x == 0) bsum[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) k_scan_bsum(int32_t *bsum, int nb, int32_t *grand) {
  __shared__ int sm[33];
  int carry = 0;
  for (int b0 = 0; b0 < nb; b0 += 1024) {
    int i = b0 + threadIdx.x;
    int v = i < nb ? bsum[i] : 0;
    int total;
    int ex = block_exclusive_scan(v, &total, sm);
    if (i < nb) bsum[i] = ex + carry;
    carry += total;
  }
  if (threadIdx.x == 0) *grand = carry;
}

__global__ void __launch_bounds__(SCAN_T) k_scan_add(int32_t *__restrict__ out, int64_t n,
                                                       const int32_t *__restrict__ bsum) {
  const int64_t base = (int64_t)blockIdx.x * SCAN_CH + (int64_t)threadIdx.x * SCAN_E;
  const int add = bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_E; k++)
    if (base + k < n) out[base + k] += add;
}

// out[0..n]: out[i] = sum(in[0..i-1]); out[n] = total.  in/out may not alias.
int exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t st) {
  if (n <= 0) {
    B2N_CUDA_OK(cudaMemsetAsync(out, 0, sizeof(int32_t), st));
    return 0;
  }
  const int nb = cdiv(n, SCAN_CH);
  int32_t *bsum = nullptr;
  if (int e = dev_alloc_t(&bsum, (size_t)nb + 1, st)) return e;
  k_scan_blocks<<<nb, SCAN_T, 0, st>>>(in, out, n, bsum);  B2N_LAUNCHED(1);
  k_scan_bsum<<<1, 1024, 0, st>>>(bsum, nb, out + n);  B2N_LAUNCHED(1);
  k_scan_add<<<nb, SCAN_T, 0, st>>>(out, n, bsum);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  dev_free(bsum, st);
  return 0;
}

// ------------------------------------------------------------------------------- bin sort
struct SortGeom {
  int dim;
  int nf[3];
  int bin[3];
  int nbin[3];
};

template <typename T>
__device__ __forceinline__ int point_bin(const SortGeom &g, T xr, T yr, T zr) {
  int b = bin_of(xr, g.bin[0], g.nbin[0]);
  if (g.dim > 1) b += g.nbin[0] * bin_of(yr, g.bin[1], g.nbin[1]);
  if (g.dim > 2) b += g.nbin[0] * g.nbin[1] * bin_of(zr, g.bin[2], g.nbin[2]);
  return b;
}

// pass 1: histogram + per-point rank inside its bin (warp-aggregated atomics)
template <typename T>
__global__ void __launch_bounds__(256) k_bin_count(SortGeom g, int64_t M, const T *__restrict__ x,
                                                    const T *__restrict__ y, const T *__restrict__ z,
                                                    int32_t *__restrict__ hist,
                                                    int32_t *__restrict__ pbin,
                                                    int32_t *__restrict__ prank) {
  const int lane = threadIdx.x & 31;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t Mpad = (M + 31) & ~int64_t(31);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < Mpad; i += stride) {
    const bool valid = i < M;
    int b = -1 - lane;  // unique dummy key for tail lanes
    if (valid) {
      T xr = fold_rescale(x[i], g.nf[0]);
      T yr = g.dim > 1 ? fold_rescale(y[i], g.nf[1]) : T(0);
      T zr = g.dim > 2 ? fold_rescale(z[i], g.nf[2]) : T(0);
      b = point_bin(g, xr, yr, zr);
    }
    const unsigned peers = __match_any_sync(0xffffffffu, b);
    const int leader = __ffs(peers) - 1;
    int base = 0;
    if (valid && lane == leader) base = atomicAdd(&hist[b], __popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (valid) {
      pbin[i] = b;
      prank[i] = base + __popc(peers & ((1u << lane) - 1u));
    }
  }
}

// pass 2: scatter index + folded coordinates to the sorted position
template <typename T>
__global__ void __launch_bounds__(256) k_bin_scatter(SortGeom g, int64_t M, const T *__restrict__ x,
                                                      const T *__restrict__ y,
                                                      const T *__restrict__ z,
                                                      const int32_t *__restrict__ bin_start,
                                                      const int32_t *__restrict__ pbin,
                                                      const int32_t *__restrict__ prank,
                                                      int32_t *__restrict__ idx, T *__restrict__ xs,
                                                      T *__restrict__ ys, T *__restrict__ zs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    const int pos = bin_start[pbin[i]] + prank[i];
    idx[pos] = (int32_t)i;
    xs[pos] = fold_rescale(x[i], g.nf[0]);
    if (g.dim > 1) ys[pos] = fold_rescale(y[i], g.nf[1]);
    if (g.dim > 2) zs[pos] = fold_rescale(z[i], g.nf[2]);
  }
}

// gpu_sort = 0: identity permutation, folded coordinates only
template <typename T>
__global__ void __launch_bounds__(256) k_fold_only(SortGeom g, int64_t M, const T *__restrict__ x,
                                                    const T *__restrict__ y, const T *__restrict__ z,
                                                    int32_t *__restrict__ idx, T *__restrict__ xs,
                                                    T *__restrict__ ys, T *__restrict__ zs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < M; i += stride) {
    idx[i] = (int32_t)i;
    xs[i] = fold_rescale(x[i], g.nf[0]);
    if (g.dim > 1) ys[i] = fold_rescale(y[i], g.nf[1]);
    if (g.dim > 2) zs[i] = fold_rescale(z[i], g.nf[2]);
  }
}

// subproblems per bin = ceil(count / maxsub)   (precision_independent.cu:75-81)
__global__ void k_sp_count(int64_t nbins, const int32_t *__restrict__ bin_start, int maxsub,
                           int32_t *__restrict__ cnt) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbins) {
    const int n = bin_start[b + 1] - bin_start[b];
    cnt[b] = (n + maxsub - 1) / maxsub;
  }
}
// subproblem -> bin map   (precision_independent.cu:83-91)
__global__ void k_sp_fill(int64_t nbins, const int32_t *__restrict__ sp_off,
                          int32_t *__restrict__ sp_bin, int64_t cap) {
  const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b < nbins) {
    const int e = sp_off[b + 1];
    for (int s = sp_off[b]; s < e && s < cap; s++) sp_bin[s] = (int32_t)b;
  }
}

template <typename T>
int binsort_points(Plan<T> &p, int64_t M, const T *x, const T *y, const T *z) {
  cudaStream_t st = p.stream;
  PointSet<T> &ps = p.pts;
  ps.M = M;
  // (re)allocate, growing only
  if (M > ps.cap_M) {
    for (int d = 0; d < 3; d++) { dev_free(ps.xs[d], st); ps.xs[d] = nullptr; }
    dev_free(ps.idx, st);
    ps.idx = nullptr;
    ps.cap_M = 0;
    for (int d = 0; d < p.dim; d++)
      if (int e = dev_alloc_t(&ps.xs[d], (size_t)M, st)) return e;
    if (int e = dev_alloc_t(&ps.idx, (size_t)M, st)) return e;
    ps.cap_M = M;
  }
  SortGeom g;
  g.dim = p.dim;
  for (int d = 0; d < 3; d++) {
    g.nf[d] = (int)p.nf[d];
    g.bin[d] = p.bin[d];
    g.nbin[d] = p.nbin[d];
  }
  const int nblk = (int)std::min<int64_t>(std::max<int64_t>(cdiv(M, 256), 1), 148 * 16);
  const bool sorted = p.opts.gpu_sort != 0 || p.method != 1;
  if (!sorted) {
    if (M > 0) k_fold_only<T><<<nblk, 256, 0, st>>>(g, M, x, y, z, ps.idx, ps.xs[0], ps.xs[1], ps.xs[2]);  B2N_LAUNCHED(1);
    B2N_LAUNCH_OK();
    ps.sp_cap = 0;
    return 0;
  }
  const int64_t nbins = p.nbins;
  if (nbins > ps.cap_bins) {
    dev_free(ps.bin_start, st);
    dev_free(ps.sp_off, st);
    ps.bin_start = ps.sp_off = nullptr;
    ps.cap_bins = 0;
    if (int e = dev_alloc_t(&ps.bin_start, (size_t)nbins + 1, st)) return e;
    if (int e = dev_alloc_t(&ps.sp_off, (size_t)nbins + 1, st)) return e;
    ps.cap_bins = nbins;
  }
  const int64_t sp_cap = std::min<int64_t>(nbins, M) + M / p.maxsub + 1;
  if (sp_cap > ps.cap_sp) {
    dev_free(ps.sp_bin, st);
    ps.sp_bin = nullptr;
    ps.cap_sp = 0;
    if (int e = dev_alloc_t(&ps.sp_bin, (size_t)sp_cap, st)) return e;
    ps.cap_sp = sp_cap;
  }
  ps.sp_cap = sp_cap;

  int32_t *hist = nullptr, *pbin = nullptr, *prank = nullptr;
  if (int e = dev_alloc_t(&hist, (size_t)nbins, st)) return e;
  if (int e = dev_alloc_t(&pbin, (size_t)M, st)) return e;
  if (int e = dev_alloc_t(&prank, (size_t)M, st)) return e;
  B2N_CUDA_OK(cudaMemsetAsync(hist, 0, sizeof(int32_t) * nbins, st));
  if (M > 0) k_bin_count<T><<<nblk, 256, 0, st>>>(g, M, x, y, z, hist, pbin, prank);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  if (int e = exclusive_scan_i32(hist, ps.bin_start, nbins, st)) return e;
  if (M > 0)
    k_bin_scatter<T><<<nblk, 256, 0, st>>>(g, M, x, y, z, ps.bin_start, pbin, prank, ps.idx,
                                           ps.xs[0], ps.xs[1], ps.xs[2]);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  // subproblem list, entirely on device (the reference reads the total back and syncs,
  // V/src/cuda/3d/spread3d_wrapper.cu:479-487)
  const int nb_blk = cdiv(nbins, 256);
  k_sp_count<<<nb_blk, 256, 0, st>>>(nbins, ps.bin_start, p.maxsub, hist);  B2N_LAUNCHED(1);
  if (int e = exclusive_scan_i32(hist, ps.sp_off, nbins, st)) return e;
  k_sp_fill<<<nb_blk, 256, 0, st>>>(nbins, ps.sp_off, ps.sp_bin, sp_cap);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  dev_free(hist, st);
  dev_free(pbin, st);
  dev_free(prank, st);
  return 0;
}

template int binsort_points<float>(Plan<float> &, int64_t, const float *, const float *, const float *);
template int binsort_points<double>(Plan<double> &, int64_t, const double *, const double *, const double *);

}  // namespace b2n
