// slab.cu -- the point exchange of the spatial multi-GPU split (jax_finufft_b200/parallel.py,
// combine="slab"; SURVEY.md 8(e): the reference has no multi-GPU code of its own, sharding is
// done above its custom call, tests/sharding_test.py:119-194).  One pass groups a rank's points
// by the rank that owns their fine-grid plane along the slowest axis and re-bases that coordinate
// to the owner's local grid of L + 2*halo planes, so that the unmodified spreader can be used
// on the receiving side; the grouped arrays then travel in four all_to_all calls with the same
// split sizes (20 bytes per point in total).
//
//   outputs, each grouped by owner (structure of arrays, so that the receiver can hand the
//   arrays to setpts / execute as they arrive): z_in[M] (in [-pi, pi) of the LOCAL grid), y[M], x[M], c[M]
//   zf  = fold_rescale(z) in float64   (common.cuh), owner = floor(zf / L)
//   z_in = (zf - owner * L + halo) * 2 pi / (L + 2 halo) - pi
//
// k_slab_count: per-CTA shared-memory histogram of the owners, one global atomic per (CTA, owner).
// k_slab_scatter: a CTA ranks its points per owner in shared memory, reserves one run per owner
// with a single global atomic and writes its points there (order inside an owner's run is free).
#include <algorithm>
#include <cstring>
#include <mutex>
#include <vector>

#include "plan.h"

namespace b2n {

constexpr int SL_T = 256, SL_E = 8, SL_MAXW = 16;

template <typename T>
__device__ __forceinline__ int slab_owner(T z, int nf0, int L, int world, double *zf_out) {
  double t = fma((double)z, 0.159154943091895345554011992339482617, 0.5);
  t -= floor(t);
  const double zf = t * (double)nf0;
  int o = (int)(zf * (1.0 / (double)L));  // any consistent choice on a slab edge is fine: the halo covers both sides
  o = o < 0 ? 0 : (o >= world ? world - 1 : o);
  *zf_out = zf;
  return o;
}

template <typename T>
__global__ void __launch_bounds__(SL_T) k_slab_count(int64_t M, const T *__restrict__ z, int nf0, int L, int world,
                                                      unsigned long long *__restrict__ counts) {
  __shared__ int cnt[SL_MAXW];
  if (threadIdx.x < SL_MAXW) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t c0 = (int64_t)blockIdx.x * (SL_T * SL_E);
#pragma unroll
  for (int e = 0; e < SL_E; e++) {
    const int64_t i = c0 + e * SL_T + threadIdx.x;
    if (i < M) {
      double zf;
      atomicAdd(&cnt[slab_owner(z[i], nf0, L, world, &zf)], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < world && cnt[threadIdx.x]) atomicAdd(&counts[threadIdx.x], (unsigned long long)cnt[threadIdx.x]);
}

template <typename T>
__global__ void __launch_bounds__(SL_T) k_slab_scatter(int64_t M, const T *__restrict__ z, const T *__restrict__ y,
                                                        const T *__restrict__ x, const cpx<T> *__restrict__ c, int nf0,
                                                        int L, int world, int halo,
                                                        const unsigned long long *__restrict__ counts,
                                                        unsigned long long *__restrict__ cursors, T *__restrict__ oz,
                                                        T *__restrict__ oy, T *__restrict__ ox,
                                                        cpx<T> *__restrict__ oc) {
  __shared__ int cnt[SL_MAXW];
  __shared__ long long base[SL_MAXW];
  if (threadIdx.x < SL_MAXW) cnt[threadIdx.x] = 0;
  __syncthreads();
  const int64_t c0 = (int64_t)blockIdx.x * (SL_T * SL_E);
  const double scale = 6.283185307179586476925286766559 / (double)(L + 2 * halo);
  int own[SL_E], rk[SL_E];
  T zin[SL_E];
#pragma unroll
  for (int e = 0; e < SL_E; e++) {
    const int64_t i = c0 + e * SL_T + threadIdx.x;
    own[e] = -1;
    if (i < M) {
      double zf;
      own[e] = slab_owner(z[i], nf0, L, world, &zf);
      zin[e] = (T)((zf - (double)own[e] * (double)L + (double)halo) * scale - 3.14159265358979323846264338327950288);
      rk[e] = atomicAdd(&cnt[own[e]], 1);
    }
  }
  __syncthreads();
  if (threadIdx.x < world) {
    long long start = 0;
    for (int o = 0; o < (int)threadIdx.x; o++) start += (long long)counts[o];
    base[threadIdx.x] = cnt[threadIdx.x]
                            ? start + (long long)atomicAdd(&cursors[threadIdx.x], (unsigned long long)cnt[threadIdx.x])
                            : 0;
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < SL_E; e++) {
    if (own[e] >= 0) {
      const int64_t i = c0 + e * SL_T + threadIdx.x;
      const long long j = base[own[e]] + rk[e];
      oz[j] = zin[e];
      oy[j] = y[i];
      ox[j] = x[i];
      oc[j] = c[i];
    }
  }
}

template <typename T>
static int slab_partition(cudaStream_t st, int64_t M, const void *p0, const void *p1, const void *p2, const void *c,
                          int64_t nf0, int world, int halo, void *o0, void *o1, void *o2, void *oc, void *counts2) {
  if (world < 1 || world > SL_MAXW || nf0 % world || M < 0) return B2N_ERR_INVALID_ARGUMENT;
  unsigned long long *cnt = (unsigned long long *)counts2;
  B2N_CUDA_OK(cudaMemsetAsync(cnt, 0, 2 * world * sizeof(unsigned long long), st));
  if (M == 0) return 0;
  const int L = (int)(nf0 / world);
  const unsigned nblk = (unsigned)cdiv(M, SL_T * SL_E);
  k_slab_count<T><<<nblk, SL_T, 0, st>>>(M, (const T *)p0, (int)nf0, L, world, cnt);
  k_slab_scatter<T><<<nblk, SL_T, 0, st>>>(M, (const T *)p0, (const T *)p1, (const T *)p2, (const cpx<T> *)c, (int)nf0, L,
                                          world, halo, cnt, cnt + world, (T *)o0, (T *)o1, (T *)o2, (cpx<T> *)oc);
  B2N_LAUNCHED(2);
  B2N_LAUNCH_OK();
  return 0;
}


// ------------------------------------------------------------------ slab / pencil FFT of the sharded type 1
// After the fine grid has been summed into z-slabs (reduce-scatter of the private grids, or the slab
// split + halo exchange), rank r holds planes [r*nzl, (r+1)*nzl) of the (nf3, nf2, nf1) grid.  The
// transform is finished without ever assembling the grid (parallel.py: slab_pencil_fft):
//   stage 1  batched 2-D cuFFT over (y, x) of the local planes, then k_slab_crop_xy: keep the
//            N1 x N2 central modes (4x less data), divide by the x and y kernel Fourier series
//            (the index maps of V/src/cuda/deconvolve_wrapper.cu:76-118) and write them grouped by the
//            rank that owns their y range: send[dst][z][y - lo(dst)][x]        -> one all_to_all
//   stage 2  the received block is (nf3, N2_local, N1), z-major: strided 1-D cuFFT along z, then
//            k_pencil_crop_z: keep N3 modes, divide by the z series -> (N3, N2_local, N1)
// cuFFT plans and the three series live in a small per-process cache keyed by the geometry.
struct ModeMap {  // mode j of n (in the requested order) -> fine-grid index and |k|
  int n, nf, modeord;
  __host__ __device__ void get(int j, int *fine, int *ak) const {
    const int k = modeord == 0 ? j - n / 2 : (j < (n + 1) / 2 ? j : j - n);
    *fine = k >= 0 ? k : nf + k;
    *ak = k >= 0 ? k : -k;
  }
};

template <typename T>
__global__ void __launch_bounds__(256) k_slab_crop_xy(const cpx<T> *__restrict__ slab, cpx<T> *__restrict__ send, int nzl,
                                                       ModeMap mx, ModeMap my, const T *__restrict__ serx,
                                                       const T *__restrict__ sery, int world) {
  const int N1 = mx.n, N2 = my.n;
  const int64_t total = (int64_t)nzl * N2 * N1;
  const int q = N2 / world, rem = N2 % world;  // shard_range: the first `rem` ranks own q + 1 rows
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j1 = (int)(i % N1);
    const int64_t t = i / N1;
    const int j2 = (int)(t % N2), z = (int)(t / N2);
    int fx, ax, fy, ay;
    mx.get(j1, &fx, &ax);
    my.get(j2, &fy, &ay);
    const cpx<T> v = slab[((int64_t)z * my.nf + fy) * mx.nf + fx];
    const T sc = T(1) / (serx[ax] * sery[ay]);
    const int d = j2 < rem * (q + 1) ? j2 / (q + 1) : rem + (q ? (j2 - rem * (q + 1)) / q : 0);
    const int lo = d * q + (d < rem ? d : rem), wd = q + (d < rem ? 1 : 0);
    // block of rank d starts after the rows of ranks < d: nzl * lo * N1 elements
    cpx<T> o;
    o.x = v.x * sc;
    o.y = v.y * sc;
    send[((int64_t)nzl * lo + (int64_t)z * wd + (j2 - lo)) * N1 + j1] = o;
  }
}

template <typename T>
__global__ void __launch_bounds__(256) k_pencil_crop_z(const cpx<T> *__restrict__ pencil, cpx<T> *__restrict__ out,
                                                        int64_t plane /* N2_local * N1 */, ModeMap mz,
                                                        const T *__restrict__ serz) {
  const int64_t total = (int64_t)mz.n * plane;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int j3 = (int)(i / plane);
    int fz, az;
    mz.get(j3, &fz, &az);
    const cpx<T> v = pencil[(int64_t)fz * plane + (i - (int64_t)j3 * plane)];
    const T sc = T(1) / serz[az];
    cpx<T> o;
    o.x = v.x * sc;
    o.y = v.y * sc;
    out[i] = o;
  }
}

namespace {
struct SlabFftKey {
  int is_double, device, kind;  // kind 0: 2-D batched (nf2, nf1) x nzl; 1: 1-D nf3, stride = batch = plane
  int64_t a, b, c;
  bool operator==(const SlabFftKey &o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; }
};
struct SlabFftEntry { SlabFftKey key; cufftHandle h; };
struct SeriesKey {
  int is_double, device, ns;
  int64_t nf;
  double beta;
  bool operator==(const SeriesKey &o) const { return std::memcmp(this, &o, sizeof(*this)) == 0; }
};
struct SeriesEntry { SeriesKey key; void *dev; };
std::mutex g_slab_mu;
std::vector<SlabFftEntry> g_slab_fft;
std::vector<SeriesEntry> g_series;

int slab_fft_plan(const SlabFftKey &k, cufftHandle *out) {
  std::lock_guard<std::mutex> lk(g_slab_mu);
  for (auto &e : g_slab_fft)
    if (e.key == k) { *out = e.h; return 0; }
  cufftHandle h;
  if (cufftCreate(&h) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
  const cufftType ft = k.is_double ? CUFFT_Z2Z : CUFFT_C2C;
  size_t work = 0;
  cufftResult r;
  if (k.kind == 0) {
    long long n[2] = {(long long)k.a, (long long)k.b};  // (nf2, nf1), contiguous planes
    r = cufftMakePlanMany64(h, 2, n, nullptr, 1, k.a * k.b, nullptr, 1, k.a * k.b, ft, k.c, &work);
  } else {
    long long n[1] = {(long long)k.a};                  // nf3 along the slowest axis
    r = cufftMakePlanMany64(h, 1, n, n, k.b, 1, n, k.b, 1, ft, k.b, &work);
  }
  if (r != CUFFT_SUCCESS) {
    cufftDestroy(h);
    fprintf(stderr, "[b200nufft] slab FFT plan failed (%d)\n", (int)r);
    return B2N_ERR_CUDA_FAILURE;
  }
  if (g_slab_fft.size() >= 16) {
    cufftDestroy(g_slab_fft.front().h);
    g_slab_fft.erase(g_slab_fft.begin());
  }
  g_slab_fft.push_back({k, h});
  *out = h;
  return 0;
}

// kernel Fourier series fwkerhalf[0 .. nf/2] on the device (hostmath.cpp: fseries_host), cached
template <typename T> int slab_series(int device, int64_t nf, int ns, double beta, cudaStream_t st, const T **out) {
  SeriesKey k;
  std::memset(&k, 0, sizeof(k));
  k.is_double = sizeof(T) == 8; k.device = device; k.ns = ns; k.nf = nf; k.beta = beta;
  std::lock_guard<std::mutex> lk(g_slab_mu);
  for (auto &e : g_series)
    if (e.key == k) { *out = (const T *)e.dev; return 0; }
  std::vector<double> h((size_t)nf / 2 + 1);
  fseries_host(nf, ns, beta, h.data());
  std::vector<T> hv(h.begin(), h.end());
  void *d = nullptr;
  B2N_CUDA_OK(cudaMalloc(&d, hv.size() * sizeof(T)));
  B2N_CUDA_OK(cudaMemcpyAsync(d, hv.data(), hv.size() * sizeof(T), cudaMemcpyHostToDevice, st));
  B2N_CUDA_OK(cudaStreamSynchronize(st));  // hv goes out of scope; once per geometry
  if (g_series.size() >= 32) {
    cudaFree(g_series.front().dev);
    g_series.erase(g_series.begin());
  }
  g_series.push_back({k, d});
  *out = (const T *)d;
  return 0;
}

template <typename T> cufftResult slab_exec(cufftHandle h, cpx<T> *d, int dir) {
  if (sizeof(T) == 4) return cufftExecC2C(h, (cufftComplex *)d, (cufftComplex *)d, dir);
  return cufftExecZ2Z(h, (cufftDoubleComplex *)d, (cufftDoubleComplex *)d, dir);
}

template <typename T>
int slab_stage1(cudaStream_t st, cpx<T> *slab, int64_t nzl, int64_t nf2, int64_t nf1, int64_t N2, int64_t N1, int world,
                int iflag, int modeord, int ns, double beta, cpx<T> *send) {
  int dev = 0;
  B2N_CUDA_OK(cudaGetDevice(&dev));
  SlabFftKey k;
  std::memset(&k, 0, sizeof(k));
  k.is_double = sizeof(T) == 8; k.device = dev; k.kind = 0; k.a = nf2; k.b = nf1; k.c = nzl;
  cufftHandle h;
  if (int e = slab_fft_plan(k, &h)) return e;
  const T *sx, *sy;
  if (int e = slab_series<T>(dev, nf1, ns, beta, st, &sx)) return e;
  if (int e = slab_series<T>(dev, nf2, ns, beta, st, &sy)) return e;
  if (cufftSetStream(h, st) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
  if (slab_exec<T>(h, slab, iflag >= 0 ? CUFFT_INVERSE : CUFFT_FORWARD) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
  const ModeMap mx = {(int)N1, (int)nf1, modeord}, my = {(int)N2, (int)nf2, modeord};
  const int64_t total = nzl * N2 * N1;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  if (total > 0) k_slab_crop_xy<T><<<grid, 256, 0, st>>>(slab, send, (int)nzl, mx, my, sx, sy, world);
  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

template <typename T>
int slab_stage2(cudaStream_t st, cpx<T> *pencil, int64_t nf3, int64_t N3, int64_t plane, int iflag, int modeord, int ns,
                double beta, cpx<T> *out) {
  int dev = 0;
  B2N_CUDA_OK(cudaGetDevice(&dev));
  if (plane <= 0 || N3 <= 0) return 0;
  SlabFftKey k;
  std::memset(&k, 0, sizeof(k));
  k.is_double = sizeof(T) == 8; k.device = dev; k.kind = 1; k.a = nf3; k.b = plane;
  cufftHandle h;
  if (int e = slab_fft_plan(k, &h)) return e;
  const T *sz;
  if (int e = slab_series<T>(dev, nf3, ns, beta, st, &sz)) return e;
  if (cufftSetStream(h, st) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
  if (slab_exec<T>(h, pencil, iflag >= 0 ? CUFFT_INVERSE : CUFFT_FORWARD) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
  const ModeMap mz = {(int)N3, (int)nf3, modeord};
  const int64_t total = N3 * plane;
  const int grid = (int)std::min<int64_t>((total + 255) / 256, 148 * 32);
  k_pencil_crop_z<T><<<grid, 256, 0, st>>>(pencil, out, plane, mz, sz);
  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}
}  // namespace

}  // namespace b2n

extern "C" int b2n_slab_partition(int is_double, void *stream, int64_t M, const void *p0, const void *p1,
                                  const void *p2, const void *c, int64_t nf0, int world, int halo, void *o0,
                                  void *o1, void *o2, void *oc, void *counts2) {
  return is_double
             ? b2n::slab_partition<double>((cudaStream_t)stream, M, p0, p1, p2, c, nf0, world, halo, o0, o1, o2, oc, counts2)
             : b2n::slab_partition<float>((cudaStream_t)stream, M, p0, p1, p2, c, nf0, world, halo, o0, o1, o2, oc, counts2);
}

extern "C" int b2n_slab_fft_xy(int is_double, void *stream, void *slab, int64_t nzl, int64_t nf2, int64_t nf1, int64_t n2,
                               int64_t n1, int world, int iflag, int modeord, int ns, double beta, void *send) {
  if (!slab || !send || nzl < 0 || world < 1 || n2 < 1 || n1 < 1 || nf2 < n2 || nf1 < n1) return B2N_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_double)
    return b2n::slab_stage1<double>(st, (double2 *)slab, nzl, nf2, nf1, n2, n1, world, iflag, modeord, ns, beta, (double2 *)send);
  return b2n::slab_stage1<float>(st, (float2 *)slab, nzl, nf2, nf1, n2, n1, world, iflag, modeord, ns, beta, (float2 *)send);
}

extern "C" int b2n_slab_fft_z(int is_double, void *stream, void *pencil, int64_t nf3, int64_t n3, int64_t plane, int iflag,
                              int modeord, int ns, double beta, void *out) {
  if (!pencil || !out || nf3 < n3 || n3 < 0 || plane < 0) return B2N_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  if (is_double) return b2n::slab_stage2<double>(st, (double2 *)pencil, nf3, n3, plane, iflag, modeord, ns, beta, (double2 *)out);
  return b2n::slab_stage2<float>(st, (float2 *)pencil, nf3, n3, plane, iflag, modeord, ns, beta, (float2 *)out);
}
