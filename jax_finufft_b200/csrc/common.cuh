// common.cuh -- shared device/host helpers for libb200nufft (sm_100a only).
#pragma once
#include <atomic>
#include <cuda_runtime.h>
#include <cufft.h>
#include <stdint.h>
#include <cstdio>
#include <type_traits>

#include "../../include/b200nufft.h"

namespace b2n {

constexpr int MAX_NS = 16;
constexpr int MIN_NS = 2;
constexpr int MAX_NCOEF = 24;
constexpr int MAX_NQUAD = 100;
constexpr double PI = 3.141592653589793238462643383279502884;

template <typename T> struct cpx_t;
template <> struct cpx_t<float> { using type = float2; };
template <> struct cpx_t<double> { using type = double2; };
template <typename T> using cpx = typename cpx_t<T>::type;

#define B2N_CUDA_OK(expr)                                                                   \
  do {                                                                                      \
    cudaError_t e__ = (expr);                                                               \
    if (e__ != cudaSuccess) {                                                               \
      fprintf(stderr, "[b200nufft] CUDA error %s at %s:%d: %s\n", cudaGetErrorName(e__),    \
              __FILE__, __LINE__, cudaGetErrorString(e__));                                 \
      return B2N_ERR_CUDA_FAILURE;                                                          \
    }                                                                                       \
  } while (0)

#define B2N_LAUNCH_OK()                                                                     \
  do {                                                                                      \
    cudaError_t e__ = cudaGetLastError();                                                   \
    if (e__ != cudaSuccess) {                                                               \
      fprintf(stderr, "[b200nufft] launch error %s at %s:%d: %s\n", cudaGetErrorName(e__),  \
              __FILE__, __LINE__, cudaGetErrorString(e__));                                 \
      return B2N_ERR_CUDA_FAILURE;                                                          \
    }                                                                                       \
  } while (0)

// Piecewise-polynomial table of the ES kernel, passed BY VALUE as a kernel parameter (lives in
// the constant bank: every access below is warp-uniform).  c[k][j]: k-th Horner coefficient
// (highest power first) of interval j.  Generated at plan time by hostmath.cpp::horner_fit --
// our own fit of the formula at V/include/cufinufft/spreadinterp.h:64-82, not the reference's
// table (V/include/cufinufft/contrib/ker_horner_allw_loop.inc).
template <typename T> struct HornerTable {
  T c[MAX_NCOEF][MAX_NS];
  int ncoef;
  int ns;
  T es_c;     // (2/ns)^2, for kerevalmeth=0
  T es_beta;  // beta,      for kerevalmeth=0
  int direct; // 1: kerevalmeth=0
};

// ---------------------------------------------------------------------------------------------
// fold_rescale: x in R -> [0,N) periodic.  Restates V/include/cufinufft/spreadinterp.h:30-57
// operation for operation (fma rn; subtract round-down; multiply round-down), because those
// roundings decide which fine-grid cell / bin a boundary point lands in.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float fold_rescale(float x, int N) {
  float r = __fmaf_rn(x, 0.159154943091895345554011992339482617f, 0.5f);
  r = __fsub_rd(r, floorf(r));
  return __fmul_rd(r, (float)N);
}
__device__ __forceinline__ double fold_rescale(double x, int N) {
  double r = __fma_rn(x, 0.159154943091895345554011992339482617, 0.5);
  r = __dsub_rd(r, floor(r));
  return __dmul_rd(r, (double)N);
}

// first fine-grid index of the ns-wide window: V/include/cufinufft/utils.h:52-56 (interval)
template <typename T> __device__ __forceinline__ int window_start(T xr, int ns) {
  return (int)ceil(xr - T(ns) * T(0.5));
}
__device__ __forceinline__ int window_start(float xr, int ns) {
  return (int)ceilf(xr - (float)ns * 0.5f);
}

// bin index of one folded coordinate: V/src/cuda/3d/spreadinterp3d.cuh:41-52
template <typename T> __device__ __forceinline__ int bin_of(T xr, int bin_size, int nbin) {
  int b = (int)floor(xr / T(bin_size));
  b = b >= nbin ? b - 1 : b;
  return b < 0 ? 0 : b;
}

__device__ __forceinline__ int wrap_once(int i, int n) {
  return i < 0 ? i + n : (i > n - 1 ? i - n : i);
}

// Kernel weights ker[j] = phi(x1 + j), j < NS, x1 = window_start - xr in [-ns/2, -ns/2+1].
// Horner form (gpu_kerevalmeth=1): z = 2 x1 + ns - 1 (same variable as spreadinterp.h:117-118).
// Direct form (gpu_kerevalmeth=0): spreadinterp.h:84-105.
template <typename T, int NS>
__device__ __forceinline__ void eval_kernel(T (&ker)[NS], T x1, const HornerTable<T> &tab) {
  if (!tab.direct) {
    const T z = fma(T(2), x1, T(NS - 1));
#pragma unroll
    for (int j = 0; j < NS; j++) ker[j] = tab.c[0][j];
    for (int k = 1; k < tab.ncoef; k++) {
#pragma unroll
      for (int j = 0; j < NS; j++) ker[j] = fma(ker[j], z, tab.c[k][j]);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NS; j++) {
      const T x = fabs(x1 + T(j));
      const T zsq = tab.es_c * x * x;
      ker[j] = (zsq < T(1)) ? exp(tab.es_beta * (sqrt(T(1) - zsq) - T(1))) : T(0);
    }
  }
}

// runtime-ns variant used by the generic (GM) kernels and by 1-D
template <typename T>
__device__ __forceinline__ void eval_kernel_rt(T *ker, int ns, T x1, const HornerTable<T> &tab) {
  if (!tab.direct) {
    const T z = fma(T(2), x1, T(ns - 1));
    for (int j = 0; j < ns; j++) {
      T v = tab.c[0][j];
      for (int k = 1; k < tab.ncoef; k++) v = fma(v, z, tab.c[k][j]);
      ker[j] = v;
    }
  } else {
    for (int j = 0; j < ns; j++) {
      const T x = fabs(x1 + T(j));
      const T zsq = tab.es_c * x * x;
      ker[j] = (zsq < T(1)) ? exp(tab.es_beta * (sqrt(T(1) - zsq) - T(1))) : T(0);
    }
  }
}

// vectorised no-return global reductions (sm_90+: red.global.add.v2/v4.f32 -> REDG F32x2/x4)
__device__ __forceinline__ void red_add(float2 *addr, float2 v) {
  asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
}
// the same, skipped (predicated, no branch) when both parts are zero
__device__ __forceinline__ void red_add_nz(float2 *addr, float2 v) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "setp.neu.f32 p, %1, 0f00000000;\n\t"
      "setp.neu.f32 q, %2, 0f00000000;\n\t"
      "or.pred p, p, q;\n\t"
      "@p red.global.add.v2.f32 [%0], {%1, %2};\n\t}" ::"l"(addr), "f"(v.x), "f"(v.y) : "memory");
}
__device__ __forceinline__ void red_add4(float4 *addr, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void red_add(double2 *addr, double2 v) {
  atomicAdd(&addr->x, v.x);
  atomicAdd(&addr->y, v.y);
}

// number of kernels of THIS library launched by the calling process (bench.py's gpu_launches)
extern std::atomic<unsigned long long> g_launch_count;  // callers may be concurrent host threads
#define B2N_LAUNCHED(n) (::b2n::g_launch_count += (n))

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

}  // namespace b2n
