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

}  // namespace b2n

extern "C" int b2n_slab_partition(int is_double, void *stream, int64_t M, const void *p0, const void *p1,
                                  const void *p2, const void *c, int64_t nf0, int world, int halo, void *o0,
                                  void *o1, void *o2, void *oc, void *counts2) {
  return is_double
             ? b2n::slab_partition<double>((cudaStream_t)stream, M, p0, p1, p2, c, nf0, world, halo, o0, o1, o2, oc, counts2)
             : b2n::slab_partition<float>((cudaStream_t)stream, M, p0, p1, p2, c, nf0, world, halo, o0, o1, o2, oc, counts2);
}
