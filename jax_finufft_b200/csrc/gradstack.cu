// gradstack.cu -- the per-point stacks and reductions of the JVP / VJP rules, one fused pass each.
// The reference's rules (src/jax_finufft/ops.py:238-273, 280-314) build the operand of their stacked
// transform as jnp.stack([c, dx*c, dy*c, dz*c], axis=2) and reduce its result with
// sum(conj(c) * h_d): a handful of elementwise XLA ops each.  Here the stack is FORMED from the
// base array and the per-dimension real scale vectors in one pass (read 8 + 4 K bytes, write
// 8 (K + 1) per point instead of K multiplies + a stack copy), and the point gradients are REDUCED
// in one pass.  Deliberately not folded into the spreader's strength load: the strengths are read
// through the sort permutation, so a scale vector read there would be a second random 128-byte
// line per point and transform.
#include "plan.h"

namespace b2n {

constexpr int GS_MAXK = 4;

template <typename T> struct StackArgs {
  const cpx<T> *src;       // [n_tot][n_transf][n]
  const T *scale[GS_MAXK]; // each [n_tot][n] (shared by the n_transf transforms) or null = 1
  cpx<T> *out;             // [n_tot][n_transf][K][n]
  int64_t n;
  int n_transf, K;
};

template <typename T> __global__ void __launch_bounds__(256) k_stack_scaled(const StackArgs<T> a) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  const int64_t it = blockIdx.y;                 // i * n_transf + t
  const int64_t i = it / a.n_transf;
  const cpx<T> v = a.src[it * a.n + j];
#pragma unroll
  for (int k = 0; k < GS_MAXK; k++) {
    if (k >= a.K) break;
    const T s = a.scale[k] ? a.scale[k][i * a.n + j] : T(1);
    cpx<T> o;
    o.x = v.x * s;
    o.y = v.y * s;
    a.out[(it * a.K + k) * a.n + j] = o;
  }
}

template <typename T> struct GradArgs {
  const cpx<T> *c;   // [n_tot][n_transf][n]
  const cpx<T> *h;   // [n_tot][n_transf][KH][n]; components k0 .. k0 + K - 1 are reduced
  T *out;            // [K][n_tot][n]
  int64_t n, n_tot;
  int n_transf, KH, k0, K, mode;  // mode 0: sign * Im(conj(c) h), 1: sign * Re(conj(c) h)
  T sign;
};

template <typename T> __global__ void __launch_bounds__(256) k_grad_points(const GradArgs<T> a) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.n) return;
  const int64_t i = blockIdx.y;
  T acc[GS_MAXK];
#pragma unroll
  for (int k = 0; k < GS_MAXK; k++) acc[k] = T(0);
  for (int t = 0; t < a.n_transf; t++) {
    const int64_t it = i * a.n_transf + t;
    const cpx<T> c = a.c[it * a.n + j];
#pragma unroll
    for (int k = 0; k < GS_MAXK; k++) {
      if (k >= a.K) break;
      const cpx<T> h = a.h[(it * a.KH + a.k0 + k) * a.n + j];
      acc[k] += a.mode == 0 ? c.x * h.y - c.y * h.x : c.x * h.x + c.y * h.y;
    }
  }
#pragma unroll
  for (int k = 0; k < GS_MAXK; k++) {
    if (k >= a.K) break;
    a.out[((int64_t)k * a.n_tot + i) * a.n + j] = a.sign * acc[k];
  }
}

template <typename T>
static int stack_scaled(cudaStream_t st, int64_t n_tot, int n_transf, int64_t n, int K, const void *src,
                        const void *const *scales, void *out) {
  StackArgs<T> a;
  a.src = (const cpx<T> *)src;
  for (int k = 0; k < GS_MAXK; k++) a.scale[k] = k < K ? (const T *)scales[k] : nullptr;
  a.out = (cpx<T> *)out;
  a.n = n;
  a.n_transf = n_transf;
  a.K = K;
  if (n > 0 && n_tot * n_transf > 0) {
    dim3 grid((unsigned)cdiv(n, 256), (unsigned)(n_tot * n_transf));
    k_stack_scaled<T><<<grid, 256, 0, st>>>(a);
    B2N_LAUNCHED(1);
  }
  B2N_LAUNCH_OK();
  return 0;
}

template <typename T>
static int grad_points(cudaStream_t st, int64_t n_tot, int n_transf, int64_t n, int KH, int k0, int K, int mode,
                       double sign, const void *c, const void *h, void *out) {
  GradArgs<T> a;
  a.c = (const cpx<T> *)c;
  a.h = (const cpx<T> *)h;
  a.out = (T *)out;
  a.n = n;
  a.n_tot = n_tot;
  a.n_transf = n_transf;
  a.KH = KH;
  a.k0 = k0;
  a.K = K;
  a.mode = mode;
  a.sign = (T)sign;
  if (n > 0 && n_tot > 0) {
    dim3 grid((unsigned)cdiv(n, 256), (unsigned)n_tot);
    k_grad_points<T><<<grid, 256, 0, st>>>(a);
    B2N_LAUNCHED(1);
  }
  B2N_LAUNCH_OK();
  return 0;
}

}  // namespace b2n

extern "C" int b2n_stack_scaled(int is_double, void *stream, int64_t n_tot, int n_transf, int64_t n, int n_scale,
                                const void *src, const void *const *scales, void *out) {
  if (n_scale < 1 || n_scale > b2n::GS_MAXK || n_tot < 0 || n_transf < 1 || n < 0 || !scales) return B2N_ERR_INVALID_ARGUMENT;
  if ((!src || !out) && n_tot * n > 0) return B2N_ERR_INVALID_ARGUMENT;
  if (n_tot * n_transf > 65535) return B2N_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  return is_double ? b2n::stack_scaled<double>(st, n_tot, n_transf, n, n_scale, src, scales, out)
                   : b2n::stack_scaled<float>(st, n_tot, n_transf, n, n_scale, src, scales, out);
}

extern "C" int b2n_grad_points(int is_double, void *stream, int64_t n_tot, int n_transf, int64_t n, int n_comp, int first,
                               int count, int mode, double sign, const void *c, const void *h, void *out) {
  if (count < 1 || count > b2n::GS_MAXK || first < 0 || first + count > n_comp || n_tot < 0 || n_tot > 65535 || n_transf < 1 ||
      n < 0 || (mode != 0 && mode != 1))
    return B2N_ERR_INVALID_ARGUMENT;
  if ((!c || !h || !out) && n_tot * n > 0) return B2N_ERR_INVALID_ARGUMENT;
  cudaStream_t st = (cudaStream_t)stream;
  return is_double ? b2n::grad_points<double>(st, n_tot, n_transf, n, n_comp, first, count, mode, sign, c, h, out)
                   : b2n::grad_points<float>(st, n_tot, n_transf, n, n_comp, first, count, mode, sign, c, h, out);
}
