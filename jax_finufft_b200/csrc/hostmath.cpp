// hostmath.cpp -- host-side plan arithmetic (float64 unless the reference rounds in float32).
// Pure C++ (no CUDA): compiled by g++ and unit-tested on CPU against oracle/ through the C ABI.
#include "hostmath.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>

namespace b2n {

// Kernel width ns and ES shape beta from the tolerance.
// Reference: V/src/cuda/spreadinterp.cpp:16-90 (setup_spreader).  For the float32 library the
// reference does this arithmetic in float (T=float: float eps, log10f as resolved by nvcc --
// SURVEY.md §0.3 -- and a float product for beta that is then widened to double).
int setup_spreader(double eps, double upsampfac, int kerevalmeth, bool is_double, int *ns_out,
                   double *beta_out) {
  int ier = 0;
  if (upsampfac != 2.0 && upsampfac != 1.25) {
    if (kerevalmeth == 1) {
      fprintf(stderr, "[b200nufft] error: nonstandard upsampfac=%.3g cannot be handled by kerevalmeth=1\n", upsampfac);
      return B2N_ERR_HORNER_WRONG_BETA;
    }
    if (upsampfac <= 1.0) {
      fprintf(stderr, "[b200nufft] error: upsampfac=%.3g\n", upsampfac);
      return B2N_ERR_UPSAMPFAC_TOO_SMALL;
    }
  }
  int ns;
  if (!is_double) {
    float e = (float)eps, s = (float)upsampfac;
    const float EPS = std::numeric_limits<float>::epsilon();
    if (e < EPS) { e = EPS; ier = B2N_WARN_EPS_TOO_SMALL; }
    ns = (int)std::ceil(-log10f(e / 10.0f));
    if (s != 2.0f) ns = (int)std::ceil(-logf(e) / ((float)PI_D * sqrtf(1.0f - 1.0f / s)));
    ns = std::max(2, ns);
    if (ns > 16) { ns = 16; ier = B2N_WARN_EPS_TOO_SMALL; }
    float b = 2.30f;
    if (ns == 2) b = 2.20f;
    if (ns == 3) b = 2.26f;
    if (ns == 4) b = 2.38f;
    if (s != 2.0f) b = 0.97f * (float)PI_D * (1.0f - 1.0f / (2.0f * s));
    *beta_out = (double)(b * (float)ns);
  } else {
    const double EPS = std::numeric_limits<double>::epsilon();
    if (eps < EPS) { eps = EPS; ier = B2N_WARN_EPS_TOO_SMALL; }
    ns = (int)std::ceil(-std::log10(eps / 10.0));
    if (upsampfac != 2.0) ns = (int)std::ceil(-std::log(eps) / (PI_D * std::sqrt(1.0 - 1.0 / upsampfac)));
    ns = std::max(2, ns);
    if (ns > 16) { ns = 16; ier = B2N_WARN_EPS_TOO_SMALL; }
    double b = 2.30;
    if (ns == 2) b = 2.20;
    if (ns == 3) b = 2.26;
    if (ns == 4) b = 2.38;
    if (upsampfac != 2.0) b = 0.97 * PI_D * (1.0 - 1.0 / (2.0 * upsampfac));
    *beta_out = b * (double)ns;
  }
  *ns_out = ns;
  return ier;
}

// smallest even integer >= n of the form 2^a 3^b 5^c that is a multiple of b.
// Reference: V/src/common/utils.cpp:124-143.
int64_t next235beven(int64_t n, int64_t b) {
  if (n <= 2) return 2;
  if (n & 1) n += 1;
  for (int64_t cand = n;; cand += 2) {
    int64_t r = cand;
    while (r % 2 == 0) r /= 2;
    while (r % 3 == 0) r /= 3;
    while (r % 5 == 0) r /= 5;
    if (r == 1 && cand % b == 0) return cand;
  }
}

// Reference: V/src/cuda/common.cu:166-177 (set_nf_type12), MAX_NF guard dropped (int32 limits
// are checked by the caller as V/src/cuda/cufinufft.cu:12-29 does).
int64_t set_nf_type12(int64_t ms, double upsampfac, int ns) {
  int64_t nf = (int64_t)std::ceil(upsampfac * (double)ms);
  if (nf < 2 * ns) nf = 2 * ns;
  return next235beven(nf, 1);
}

// P_n(x) and P_n'(x) by the three-term recurrence.  Reference: V/src/common/utils.cpp:66-86.
static void legendre(int n, double x, double &p, double &dp) {
  if (n == 0) { p = 1.0; dp = 0.0; return; }
  if (n == 1) { p = x; dp = 1.0; return; }
  double a = 1.0, b = x;
  for (int i = 1; i < n; i++) {
    double c = ((2 * i + 1) * x * b - i * a) / (i + 1);
    a = b;
    b = c;
  }
  p = b;
  dp = n * (x * b - a) / (x * x - 1);
}

// n-node Gauss-Legendre rule on [-1,1] (Newton from Chebyshev guesses, stop after the step is
// below 1e-14 three times).  Reference: V/src/common/utils.cpp:25-64.
void gaussquad(int n, double *x, double *w) {
  x[n / 2] = 0.0;
  for (int i = 0; i < n / 2; i++) {
    double t = std::cos((2 * i + 1) * PI_D / (2 * n));
    int hits = 0;
    while (hits < 3) {
      double p, dp;
      legendre(n, t, p, dp);
      double dt = -p / dp;
      t += dt;
      if (std::fabs(dt) < 1e-14) hits++;
    }
    x[i] = -t;
    x[n - 1 - i] = t;
  }
  for (int i = 0; i <= n / 2; i++) {
    double p0, dp, p1, d1;
    legendre(n, x[i], p0, dp);
    legendre(n + 1, x[i], p1, d1);
    w[i] = w[n - 1 - i] = -2.0 / ((n + 1) * dp * p1);
  }
}

// exact ES kernel; reference: V/include/cufinufft/spreadinterp.h:64-82
double es_kernel(double x, int ns, double beta) {
  double z = 2.0 * x / (double)ns;
  if (std::fabs(z) >= 1.0) return 0.0;
  return std::exp(beta * (std::sqrt(1.0 - z * z) - 1.0));
}

// Quadrature nodes for the kernel Fourier transform.
//  type 1/2 (equispaced):  q = floor(2 + 1.5 ns) nodes, f_n = (ns/2) w_n phi(z_n),
//     phase_n = 2 pi z_n / nf      -- V/src/cuda/common.cu:196-209
//  type 3 (arbitrary freq): q = floor(2 + ns) nodes, z_n, f_n   -- V/src/cuda/common.cu:211-224
int kernel_quadrature(int ns, double beta, bool type3, int64_t nf, double *f, double *zp) {
  double J2 = ns / 2.0;
  int q = type3 ? (int)(2 + 2.0 * J2) : (int)(2 + 3.0 * J2);
  std::vector<double> z(2 * q), w(2 * q);
  gaussquad(2 * q, z.data(), w.data());
  for (int n = 0; n < q; n++) {
    double zn = z[n] * J2;
    f[n] = J2 * w[n] * es_kernel(zn, ns, beta);
    zp[n] = type3 ? zn : 2.0 * PI_D * zn / (double)nf;
  }
  return q;
}

// host evaluation of fwkerhalf (used by CPU tests; the plan computes it on the GPU)
void fseries_host(int64_t nf, int ns, double beta, double *out) {
  double f[MAX_NQUAD_H], ph[MAX_NQUAD_H];
  int q = kernel_quadrature(ns, beta, false, nf, f, ph);
  for (int64_t i = 0; i <= nf / 2; i++) {
    double s = 0.0;
    for (int n = 0; n < q; n++) s += f[n] * 2.0 * std::cos((double)i * ph[n]);
    out[i] = (i & 1) ? -s : s;
  }
}

// Reference: V/include/cufinufft/utils.h:154-183 (set_nhg_type3).  The reference evaluates this
// in the library precision T; `is_double=false` reproduces the float arithmetic.
template <typename T> static void nhg3(T S, T X, double upsampfac, int ns, int64_t *nf, double *h, double *gam) {
  int nss = ns + 1;
  T Xsafe = X, Ssafe = S;
  if (X == 0.0) {
    if (S == 0.0) { Xsafe = 1.0; Ssafe = 1.0; }
    else Xsafe = std::max(Xsafe, T(1) / S);
  } else
    Ssafe = std::max(Ssafe, T(1) / X);
  T nfd = (T)(2.0 * upsampfac * Ssafe * Xsafe / PI_D + nss);
  if (!std::isfinite(nfd)) nfd = 0.0;
  int64_t n = (int64_t)(int)nfd;
  if (n < 2 * ns) n = 2 * ns;
  n = next235beven(n, 1);
  *nf = n;
  *h = (double)(2 * T(PI_D) / n);
  *gam = (double)(T)(T(n) / (2.0 * upsampfac * Ssafe));
}
void set_nhg_type3(double S, double X, double upsampfac, int ns, bool is_double, int64_t *nf, double *h, double *gam) {
  if (is_double) nhg3<double>(S, X, upsampfac, ns, nf, h, gam);
  else nhg3<float>((float)S, (float)X, upsampfac, ns, nf, h, gam);
}

// Reference: V/include/cufinufft/utils.h:143-152 (arraywidcen), GROWFRAC = 0.1
void widcen(double lo, double hi, bool is_double, double *w, double *c) {
  if (is_double) {
    double ww = (hi - lo) / 2, cc = (hi + lo) / 2;
    if (std::fabs(cc) < 0.1 * ww) { ww += std::fabs(cc); cc = 0.0; }
    *w = ww; *c = cc;
  } else {
    float l = (float)lo, hh = (float)hi;
    float ww = (hh - l) / 2, cc = (hh + l) / 2;
    if (std::fabs(cc) < 0.1 * ww) { ww += std::fabs(cc); cc = 0.0f; }
    *w = ww; *c = cc;
  }
}

// ---------------------------------------------------------------------------------------------
// Piecewise-polynomial ("Horner") table of the ES kernel -- our own fit, generated at plan time.
// For interval j (weight j of the ns-wide window) the kernel argument is
//     x = (z - ns + 1)/2 + j,   z in [-1,1]
// and ker_j(z) = phi(x).  We interpolate at nc Chebyshev nodes and convert to monomials in z by
// a pivoted dense solve (nc <= 24; the monomial basis on [-1,1] is benign at these degrees).
// nc is the smallest count whose max error (relative to the kernel peak 1) is below `tol`.
// ---------------------------------------------------------------------------------------------
static bool solve_dense(int n, std::vector<double> &A, std::vector<double> &b) {
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (std::fabs(A[r * n + c]) > std::fabs(A[piv * n + c])) piv = r;
    if (A[piv * n + c] == 0.0) return false;
    if (piv != c) {
      for (int k = 0; k < n; k++) std::swap(A[c * n + k], A[piv * n + k]);
      std::swap(b[c], b[piv]);
    }
    for (int r = c + 1; r < n; r++) {
      double m = A[r * n + c] / A[c * n + c];
      if (m == 0.0) continue;
      for (int k = c; k < n; k++) A[r * n + k] -= m * A[c * n + k];
      b[r] -= m * b[c];
    }
  }
  for (int r = n - 1; r >= 0; r--) {
    double s = b[r];
    for (int k = r + 1; k < n; k++) s -= A[r * n + k] * b[k];
    b[r] = s / A[r * n + r];
  }
  return true;
}

static double fit_interval(int ns, double beta, int j, int nc, double *coef_hi_first) {
  std::vector<double> A(nc * nc), b(nc);
  for (int i = 0; i < nc; i++) {
    double z = std::cos(PI_D * (i + 0.5) / nc);
    double p = 1.0;
    for (int k = 0; k < nc; k++) { A[i * nc + k] = p; p *= z; }
    b[i] = es_kernel(0.5 * (z - ns + 1) + j, ns, beta);
  }
  solve_dense(nc, A, b);  // b[k] = coefficient of z^k
  for (int k = 0; k < nc; k++) coef_hi_first[k] = b[nc - 1 - k];
  double err = 0.0;
  const int NT = 400;
  for (int t = 0; t <= NT; t++) {
    double z = -1.0 + 2.0 * t / NT;
    double v = 0.0;
    for (int k = 0; k < nc; k++) v = v * z + coef_hi_first[k];
    err = std::max(err, std::fabs(v - es_kernel(0.5 * (z - ns + 1) + j, ns, beta)));
  }
  return err;
}

int horner_fit(int ns, double beta, bool is_double, double *coef /* [MAX_NCOEF_H][16] */) {
  // The ES kernel has a square-root endpoint singularity at |x| = ns/2, so the fit error of the
  // two outermost intervals saturates (at ~0.15 * 10^(1-ns)) however many coefficients are
  // used.  Take the smallest count that is within 30% of that floor (or already below the
  // arithmetic's own resolution): ns+2 at sigma=2, the degree the reference tables also use
  // (V/include/cufinufft/contrib/ker_horner_allw_loop.inc), instead of paying for 24.
  const double floor_tol = is_double ? 2e-15 : 1.5e-8;
  double tmp[MAX_NCOEF_H];
  double err[MAX_NCOEF_H + 1];
  for (int nc = 4; nc <= MAX_NCOEF_H; nc++) {
    err[nc] = 0.0;
    for (int j = 0; j < ns; j++) err[nc] = std::max(err[nc], fit_interval(ns, beta, j, nc, tmp));
  }
  double best = err[4];
  for (int nc = 5; nc <= MAX_NCOEF_H; nc++) best = std::min(best, err[nc]);
  int nc_used = MAX_NCOEF_H;
  for (int nc = 4; nc <= MAX_NCOEF_H; nc++)
    if (err[nc] <= std::max(1.3 * best, floor_tol)) { nc_used = nc; break; }
  std::memset(coef, 0, sizeof(double) * MAX_NCOEF_H * 16);
  for (int j = 0; j < ns; j++) {
    fit_interval(ns, beta, j, nc_used, tmp);
    for (int k = 0; k < nc_used; k++) coef[k * 16 + j] = tmp[k];
  }
  return nc_used;
}

}  // namespace b2n
