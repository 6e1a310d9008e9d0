/*
 * nufft_oracle.c -- CPU restatement of the reference GPU NUFFT algorithm (cuFINUFFT as
 * driven by jax-finufft).  TEST INFRASTRUCTURE ONLY: imported by tests/, by
 * __graft_entry__.smoke() and by bench.py's cpu_baseline / --impl reference legs as the
 * checker / CPU baseline.  The product (jax_finufft_b200) never links, loads or calls it.
 *
 * Everything is float64 and follows the reference's *algorithm* (same ns/beta/nf rules, the
 * exact exponential-of-semicircle kernel the reference's Horner tables were fitted to, the same
 * fold/rescale + window geometry, the same Gauss-Legendre kernel-FT quadrature, the same
 * deconvolve index maps, the same type-3 rescaling).  `prec==1` additionally emulates the
 * float32 roundings that decide *where* a point lands (fold_rescale with its directed
 * roundings), so that a float32 GPU run and this oracle agree on every cell/bin index.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * /root/reference; V/ = vendor/finufft/).
 *
 * Parity pinning: the reference stores no golden vectors; every reference
 * test checks against an on-the-fly direct NUDFT with fixed seeds.  This oracle is pinned the
 * same way (tests/test_oracle.py: seeds/sizes/tolerances of tests/ops_test.py:25-126 and
 * V/test/cuda/cufinufft3d_test.cu:186-254), against outputs of the unmodified reference cuFINUFFT
 * committed as tests/golden/ref_cufinufft_golden.npz (tests/test_oracle.py), and, on the GPU box,
 * against the live library built into oracle/_ref (tests/test_gpu_fullsize_vs_reference.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_PI 3.141592653589793238462643383279502884
#define MAX_NSPREAD 16
#define MIN_NSPREAD 2
#define MAX_NQUAD 100

typedef struct { double re, im; } cpx;

/* ------------------------------------------------------------------ plan arithmetic */

/* V/src/cuda/spreadinterp.cpp:48-90 (setup_spreader).  is_float selects the float32 code path
 * the reference takes for cufinufftf_* (float eps, float log10 -- SURVEY.md §0.3).
 * Returns ns; *beta_out gets the ES shape parameter; *ier gets 0 / 1 (eps too small) / 7 / 8. */
int orc_setup_spreader(double eps, double upsampfac, int kerevalmeth, int is_float,
                       double *beta_out, int *ier) {
  *ier = 0;
  if (upsampfac != 2.0 && upsampfac != 1.25) {
    if (kerevalmeth == 1) { *ier = 8; return 0; }       /* FINUFFT_ERR_HORNER_WRONG_BETA */
    if (upsampfac <= 1.0) { *ier = 7; return 0; }       /* FINUFFT_ERR_UPSAMPFAC_TOO_SMALL */
  }
  int ns;
  double betaoverns;
  if (is_float) {
    float e = (float)eps, s = (float)upsampfac;
    const float EPSF = 1.1920928955078125e-07f;
    if (e < EPSF) { e = EPSF; *ier = 1; }
    ns = (int)ceilf(-log10f(e / 10.0f));
    if (s != 2.0f) ns = (int)ceilf(-logf(e) / ((float)ORC_PI * sqrtf(1.0f - 1.0f / s)));
    if (ns < 2) ns = 2;
    if (ns > MAX_NSPREAD) { ns = MAX_NSPREAD; *ier = 1; }
    float b = 2.30f;
    if (ns == 2) b = 2.20f;
    if (ns == 3) b = 2.26f;
    if (ns == 4) b = 2.38f;
    if (s != 2.0f) b = 0.97f * (float)ORC_PI * (1.0f - 1.0f / (2.0f * s));
    /* spopts.beta is a double field: the float product is widened on assignment */
    *beta_out = (double)(b * (float)ns);
    return ns;
  }
  const double EPSD = 2.220446049250313e-16;
  if (eps < EPSD) { eps = EPSD; *ier = 1; }
  ns = (int)ceil(-log10(eps / 10.0));
  if (upsampfac != 2.0) ns = (int)ceil(-log(eps) / (ORC_PI * sqrt(1.0 - 1.0 / upsampfac)));
  if (ns < 2) ns = 2;
  if (ns > MAX_NSPREAD) { ns = MAX_NSPREAD; *ier = 1; }
  betaoverns = 2.30;
  if (ns == 2) betaoverns = 2.20;
  if (ns == 3) betaoverns = 2.26;
  if (ns == 4) betaoverns = 2.38;
  if (upsampfac != 2.0) betaoverns = 0.97 * ORC_PI * (1.0 - 1.0 / (2.0 * upsampfac));
  *beta_out = betaoverns * (double)ns;
  return ns;
}

/* V/src/common/utils.cpp:124-143 (next235beven) */
long orc_next235beven(long n, long b) {
  if (n <= 2) return 2;
  if (n % 2 == 1) n += 1;
  long nplus = n - 2, numdiv = 2;
  while (numdiv > 1 || nplus % b != 0) {
    nplus += 2;
    numdiv = nplus;
    while (numdiv % 2 == 0) numdiv /= 2;
    while (numdiv % 3 == 0) numdiv /= 3;
    while (numdiv % 5 == 0) numdiv /= 5;
  }
  return nplus;
}

/* V/src/cuda/common.cu:166-177 (set_nf_type12) */
long orc_set_nf_type12(long ms, double upsampfac, int ns) {
  long nf = (long)ceil(upsampfac * (double)ms);
  if (nf < 2 * ns) nf = 2 * ns;
  return orc_next235beven(nf, 1);
}

/* V/src/common/utils.cpp:66-86 (leg_eval) */
static void leg_eval(int n, double x, double *p, double *dp) {
  if (n == 0) { *p = 1.0; *dp = 0.0; return; }
  if (n == 1) { *p = x; *dp = 1.0; return; }
  double p0 = 0.0, p1 = 1.0, p2 = x;
  for (int i = 1; i < n; i++) {
    p0 = p1; p1 = p2;
    p2 = ((2 * i + 1) * x * p1 - i * p0) / (i + 1);
  }
  *p = p2;
  *dp = n * (x * p2 - p1) / (x * x - 1);
}

/* V/src/common/utils.cpp:25-64 (gaussquad): n-node Gauss-Legendre by Newton from Chebyshev */
void orc_gaussquad(int n, double *xgl, double *wgl) {
  xgl[n / 2] = 0;
  for (int i = 0; i < n / 2; i++) {
    int conv = 0;
    double x = cos((2 * i + 1) * ORC_PI / (2 * n));
    for (;;) {
      double p, dp;
      leg_eval(n, x, &p, &dp);
      double dx = -p / dp;
      x += dx;
      if (fabs(dx) < 1e-14) conv++;
      if (conv == 3) break;
    }
    xgl[i] = -x;
    xgl[n - i - 1] = x;
  }
  for (int i = 0; i < n / 2 + 1; i++) {
    double j1, dp, p, j2;
    leg_eval(n, xgl[i], &j1, &dp);
    leg_eval(n + 1, xgl[i], &p, &j2);
    wgl[i] = -2 / ((n + 1) * dp * p);
    wgl[n - i - 1] = wgl[i];
  }
}

/* V/include/cufinufft/spreadinterp.h:64-82 (evaluate_kernel): phi(x)=exp(beta(sqrt(1-(2x/ns)^2)-1)) */
double orc_es_kernel(double x, int ns, double beta) {
  double z = 2.0 * x / (double)ns;
  if (fabs(z) >= 1.0) return 0.0;
  return exp(beta * (sqrt(1.0 - z * z) - 1.0));
}

/* V/src/cuda/common.cu:196-209 (onedim_fseries_kernel_precomp) + 27-60
 * (cu_fseries_kernel_compute): fwkerhalf[i], i=0..nf/2 */
void orc_fseries(long nf, int ns, double beta, double *fwkerhalf) {
  double J2 = ns / 2.0;
  int q = (int)(2 + 3.0 * J2);
  double z[2 * MAX_NQUAD], w[2 * MAX_NQUAD], f[MAX_NQUAD], ph[MAX_NQUAD];
  orc_gaussquad(2 * q, z, w);
  for (int n = 0; n < q; n++) {
    z[n] *= J2;
    f[n] = J2 * w[n] * orc_es_kernel(z[n], ns, beta);
    ph[n] = 2.0 * ORC_PI * z[n] / (double)nf;
  }
  for (long i = 0; i <= nf / 2; i++) {
    double x = 0.0;
    for (int n = 0; n < q; n++) x += f[n] * 2.0 * cos((double)i * ph[n]);
    fwkerhalf[i] = x * ((i % 2) ? -1.0 : 1.0);
  }
}

/* V/src/cuda/common.cu:211-224 (onedim_nuft_kernel_precomp) + 68-102 (cu_nuft_kernel_compute):
 * kernel FT at N arbitrary frequencies k (in "grid" radians) */
void orc_nuft(int ns, double beta, long N, const double *k, double *phihat) {
  double J2 = ns / 2.0;
  int q = (int)(2 + 2.0 * J2);
  double z[2 * MAX_NQUAD], w[2 * MAX_NQUAD], f[MAX_NQUAD];
  orc_gaussquad(2 * q, z, w);
  for (int n = 0; n < q; n++) {
    z[n] *= J2;
    f[n] = J2 * w[n] * orc_es_kernel(z[n], ns, beta);
  }
#pragma omp parallel for schedule(static)
  for (long i = 0; i < N; i++) {
    double x = 0.0;
    for (int n = 0; n < q; n++) x += f[n] * 2.0 * cos(k[i] * z[n]);
    phihat[i] = x;
  }
}

/* ------------------------------------------------------------------ point geometry */

static float round_down_to_float(double v) {
  float f = (float)v;
  if ((double)f > v) f = nextafterf(f, -INFINITY);
  return f;
}

/* V/include/cufinufft/spreadinterp.h:30-57 (fold_rescale), device branch.
 * prec==1: float32 with fma(rn), subtract(rd), multiply(rd) exactly as __fmaf_rn/__fsub_rd/
 * __fmul_rd do (each float op's exact result fits a double, which is then rounded down).
 * prec==0: the same three operations in float64 (directed rounding emulated with nextafter
 * on the rare inexact results is unnecessary for float64 test tolerances: round-to-nearest). */
double orc_fold_rescale(double x, long N, int prec) {
  if (prec == 1) {
    const float x2pi = 0.159154943091895345554011992339482617f;
    float r = fmaf((float)x, x2pi, 0.5f);
    float fl = floorf(r);
    float d = round_down_to_float((double)r - (double)fl);
    return (double)round_down_to_float((double)d * (double)(float)N);
  }
  const double x2pi = 0.159154943091895345554011992339482617;
  double r = fma(x, x2pi, 0.5);
  r = r - floor(r);
  double v = r * (double)N;
  if (v >= (double)N) v = nextafter((double)N, 0.0); /* what the round-down multiply guarantees */
  return v;
}

/* V/include/cufinufft/utils.h:52-56 (interval): first grid index of the ns-wide window */
static inline long window_start(double xr, int ns) { return (long)ceil(xr - 0.5 * (double)ns); }

/* V/src/cuda/3d/spreadinterp3d.cuh:38-52 (and 2d:76-88, 1d:58-62): bin index of one coordinate */
static inline int bin_of(double xr, int bin_size, int nbin) {
  int b = (int)floor(xr / (double)bin_size);
  if (b >= nbin) b -= 1;
  if (b < 0) b = 0;
  return b;
}

/* V/src/cuda/3d/spreadinterp3d.cuh:28-56 (calc_bin_size_noghost_3d) restated without the
 * atomics: writes each point's bin id (x fastest) and the per-bin histogram.  The reference's
 * order *inside* a bin is atomicAdd-arrival order, i.e. unspecified; the
 * checkable contract is bin membership + histogram + exclusive-scan offsets. */
void orc_binsort(int dim, long M, const double *x, const double *y, const double *z, long nf1,
                 long nf2, long nf3, int bx, int by, int bz, int prec, int32_t *binid,
                 int32_t *hist /* nbinx*nbiny*nbinz, zeroed here */) {
  int nbx = (int)((nf1 + bx - 1) / bx);
  int nby = dim > 1 ? (int)((nf2 + by - 1) / by) : 1;
  int nbz = dim > 2 ? (int)((nf3 + bz - 1) / bz) : 1;
  memset(hist, 0, sizeof(int32_t) * (size_t)nbx * nby * nbz);
  for (long i = 0; i < M; i++) {
    int b = bin_of(orc_fold_rescale(x[i], nf1, prec), bx, nbx);
    if (dim > 1) b += nbx * bin_of(orc_fold_rescale(y[i], nf2, prec), by, nby);
    if (dim > 2) b += nbx * nby * bin_of(orc_fold_rescale(z[i], nf3, prec), bz, nbz);
    binid[i] = b;
    hist[b]++;
  }
}

/* Anchor-cell binning used by OUR 3-D sliding-window kernels (jax_finufft_b200/csrc/
 * swr_kernels.cuh; not a reference function): the bin of a point is taken from the cell
 * u = window_start(x') + ns/2 (periodic: u == nf -> 0), i.e. from the window of
 * V/include/cufinufft/utils.h:52-56 (interval) rather than from floor(x').  Also returns the
 * anchor z cell so that the test can check the in-bin z ordering. */
void orc_binsort_anchor(long M, const double *x, const double *y, const double *z, long nf1, long nf2,
                        long nf3, int bx, int by, int bz, int ns, int prec, int32_t *binid,
                        int32_t *hist, int32_t *uz_out) {
  int nbx = (int)((nf1 + bx - 1) / bx), nby = (int)((nf2 + by - 1) / by), nbz = (int)((nf3 + bz - 1) / bz);
  const double *p[3] = {x, y, z};
  long nf[3] = {nf1, nf2, nf3};
  memset(hist, 0, sizeof(int32_t) * (size_t)nbx * nby * nbz);
  for (long i = 0; i < M; i++) {
    long u[3];
    for (int d = 0; d < 3; d++) {
      u[d] = window_start(orc_fold_rescale(p[d][i], nf[d], prec), ns) + ns / 2;
      if (u[d] >= nf[d]) u[d] -= nf[d];
    }
    int b = (int)(u[0] / bx) + nbx * ((int)(u[1] / by) + nby * (int)(u[2] / bz));
    binid[i] = b;
    uz_out[i] = (int32_t)u[2];
    hist[b]++;
  }
}

static inline void es_weights(double xr, int ns, double beta, long *start, double *ker) {
  long xs = window_start(xr, ns);
  double x1 = (double)xs - xr;
  for (int j = 0; j < ns; j++) ker[j] = orc_es_kernel(fabs(x1 + j), ns, beta);
  *start = xs;
}

static inline long wrap_idx(long i, long n) { return i < 0 ? i + n : (i > n - 1 ? i - n : i); }

/* Type-1 gridding.  Restates V/src/cuda/3d/spreadinterp3d.cuh:86-135 (spread_3d_nupts_driven;
 * 2d:116-160, 1d twins) with the exact ES kernel (gpu_kerevalmeth=0 formula,
 * V/include/cufinufft/spreadinterp.h:84-105).  fw is ADDED to (caller zeroes).  Layout
 * fw[ix + iy*nf1 + iz*nf1*nf2].  OpenMP: threads own contiguous slabs of the slowest dim, so
 * no atomics are needed and the result is deterministic. */
void orc_spread(int dim, long M, const double *x, const double *y, const double *z, const cpx *c,
                long nf1, long nf2, long nf3, int ns, double beta, int prec, cpx *fw) {
  if (dim < 2) nf2 = 1;
  if (dim < 3) nf3 = 1;
  long nslow = dim == 3 ? nf3 : (dim == 2 ? nf2 : nf1);
  const double *slow = dim == 3 ? z : (dim == 2 ? y : x);
  /* bucket the points by slowest-dim window start (counting sort), so that each slab owner
   * only visits the points whose window can touch its slab */
  long nb = nslow + ns + 2;
  long *start = (long *)calloc((size_t)nb + 1, sizeof(long));
  int32_t *cell = (int32_t *)malloc(sizeof(int32_t) * (size_t)(M > 0 ? M : 1));
  long *order = (long *)malloc(sizeof(long) * (size_t)(M > 0 ? M : 1));
  for (long i = 0; i < M; i++) {
    long s = window_start(orc_fold_rescale(slow[i], nslow, prec), ns) + ns; /* >= 0 */
    cell[i] = (int32_t)s;
    start[s + 1]++;
  }
  for (long b = 0; b < nb; b++) start[b + 1] += start[b];
  {
    long *fill = (long *)malloc(sizeof(long) * (size_t)nb);
    memcpy(fill, start, sizeof(long) * (size_t)nb);
    for (long i = 0; i < M; i++) order[fill[cell[i]]++] = i;
    free(fill);
  }
#pragma omp parallel
  {
    int nt = 1, tid = 0;
#ifdef _OPENMP
    nt = omp_get_num_threads();
    tid = omp_get_thread_num();
#endif
    long lo = nslow * tid / nt, hi = nslow * (tid + 1) / nt; /* owned slab [lo,hi) */
    double k1[MAX_NSPREAD], k2[MAX_NSPREAD], k3[MAX_NSPREAD];
    /* visit every bucket (window start s = b - ns) whose ns-wide window, wrapped once, touches
     * the owned slab; each point is then visited exactly once per owning thread */
    for (long b = 0; b < nb; b++) {
      if (start[b] == start[b + 1]) continue;
      int hit = 0;
      for (int j = 0; j < ns && !hit; j++) {
        long iw = wrap_idx(b - ns + j, nslow);
        hit = (iw >= lo && iw < hi);
      }
      if (!hit) continue;
      for (long q = start[b]; q < start[b + 1]; q++) {
        long i = order[q];
        long xs = 0, ys = 0, zs = 0;
        es_weights(orc_fold_rescale(x[i], nf1, prec), ns, beta, &xs, k1);
        if (dim > 1) es_weights(orc_fold_rescale(y[i], nf2, prec), ns, beta, &ys, k2);
        if (dim > 2) es_weights(orc_fold_rescale(z[i], nf3, prec), ns, beta, &zs, k3);
        if (dim == 1) {
          for (int a = 0; a < ns; a++) {
            long ix = wrap_idx(xs + a, nf1);
            if (ix < lo || ix >= hi) continue;
            fw[ix].re += c[i].re * k1[a];
            fw[ix].im += c[i].im * k1[a];
          }
        } else if (dim == 2) {
          for (int bq = 0; bq < ns; bq++) {
            long iy = wrap_idx(ys + bq, nf2);
            if (iy < lo || iy >= hi) continue;
            for (int a = 0; a < ns; a++) {
              long ix = wrap_idx(xs + a, nf1);
              double w = k1[a] * k2[bq];
              cpx *o = &fw[ix + iy * nf1];
              o->re += c[i].re * w;
              o->im += c[i].im * w;
            }
          }
        } else {
          for (int cq = 0; cq < ns; cq++) {
            long iz = wrap_idx(zs + cq, nf3);
            if (iz < lo || iz >= hi) continue;
            for (int bq = 0; bq < ns; bq++) {
              long iy = wrap_idx(ys + bq, nf2);
              double w23 = k2[bq] * k3[cq];
              for (int a = 0; a < ns; a++) {
                long ix = wrap_idx(xs + a, nf1);
                double w = k1[a] * w23;
                cpx *o = &fw[ix + iy * nf1 + iz * nf1 * nf2];
                o->re += c[i].re * w;
                o->im += c[i].im * w;
              }
            }
          }
        }
      }
    }
  }
  free(start); free(cell); free(order);
}

/* Type-2 gather.  Restates V/src/cuda/3d/spreadinterp3d.cuh:556-608 (interp_3d_nupts_driven;
 * 2d:327-366, 1d twins). */
void orc_interp(int dim, long M, const double *x, const double *y, const double *z, cpx *c,
                long nf1, long nf2, long nf3, int ns, double beta, int prec, const cpx *fw) {
  if (dim < 2) nf2 = 1;
  if (dim < 3) nf3 = 1;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < M; i++) {
    double k1[MAX_NSPREAD], k2[MAX_NSPREAD], k3[MAX_NSPREAD];
    long xs = 0, ys = 0, zs = 0;
    es_weights(orc_fold_rescale(x[i], nf1, prec), ns, beta, &xs, k1);
    if (dim > 1) es_weights(orc_fold_rescale(y[i], nf2, prec), ns, beta, &ys, k2); else { k2[0] = 1; }
    if (dim > 2) es_weights(orc_fold_rescale(z[i], nf3, prec), ns, beta, &zs, k3); else { k3[0] = 1; }
    int n2 = dim > 1 ? ns : 1, n3 = dim > 2 ? ns : 1;
    double sr = 0, si = 0;
    for (int cq = 0; cq < n3; cq++) {
      long iz = dim > 2 ? wrap_idx(zs + cq, nf3) : 0;
      for (int bq = 0; bq < n2; bq++) {
        long iy = dim > 1 ? wrap_idx(ys + bq, nf2) : 0;
        double w23 = k2[bq] * k3[cq];
        for (int a = 0; a < ns; a++) {
          long ix = wrap_idx(xs + a, nf1);
          double w = k1[a] * w23;
          const cpx *g = &fw[ix + iy * nf1 + iz * nf1 * nf2];
          sr += g->re * w;
          si += g->im * w;
        }
      }
    }
    c[i].re = sr;
    c[i].im = si;
  }
}

/* ------------------------------------------------------------------ deconvolve / amplify */

/* V/src/cuda/deconvolve_wrapper.cu:16-118: index maps of deconvolve_{1,2,3}d, both modeords */
static inline void mode_map(long k, long ms, long nf, int modeord, long *w, long *kerind) {
  if (modeord == 0) {
    long p = k - ms / 2;
    *w = p >= 0 ? p : nf + p;
    *kerind = labs(p);
  } else {
    long p = k - ms + ms / 2;
    *w = p >= 0 ? nf + k - ms : k;
    *kerind = p >= 0 ? ms - k : k;
  }
}

/* deconvolve (type 1 step 3): fk[k] = fw[w(k)] / prod fwkerhalf_d[|k_d|] */
void orc_deconvolve(int dim, long ms, long mt, long mu, long nf1, long nf2, long nf3,
                    const cpx *fw, cpx *fk, const double *h1, const double *h2,
                    const double *h3, int modeord) {
  if (dim < 2) { mt = 1; nf2 = 1; }
  if (dim < 3) { mu = 1; nf3 = 1; }
#pragma omp parallel for schedule(static)
  for (long k3 = 0; k3 < mu; k3++) {
    long w3 = 0, i3 = 0;
    if (dim > 2) mode_map(k3, mu, nf3, modeord, &w3, &i3);
    for (long k2 = 0; k2 < mt; k2++) {
      long w2 = 0, i2 = 0;
      if (dim > 1) mode_map(k2, mt, nf2, modeord, &w2, &i2);
      for (long k1 = 0; k1 < ms; k1++) {
        long w1, i1;
        mode_map(k1, ms, nf1, modeord, &w1, &i1);
        double kv = h1[i1];
        if (dim > 1) kv *= h2[i2];
        if (dim > 2) kv *= h3[i3];
        const cpx *g = &fw[w1 + w2 * nf1 + w3 * nf1 * nf2];
        cpx *o = &fk[k1 + k2 * ms + k3 * ms * mt];
        o->re = g->re / kv;
        o->im = g->im / kv;
      }
    }
  }
}

/* V/src/cuda/deconvolve_wrapper.cu:121-223,316-325: amplify_{1,2,3}d after zeroing all of fw */
void orc_amplify(int dim, long ms, long mt, long mu, long nf1, long nf2, long nf3, cpx *fw,
                 const cpx *fk, const double *h1, const double *h2, const double *h3,
                 int modeord) {
  if (dim < 2) { mt = 1; nf2 = 1; }
  if (dim < 3) { mu = 1; nf3 = 1; }
  memset(fw, 0, sizeof(cpx) * (size_t)(nf1 * nf2 * nf3));
#pragma omp parallel for schedule(static)
  for (long k3 = 0; k3 < mu; k3++) {
    long w3 = 0, i3 = 0;
    if (dim > 2) mode_map(k3, mu, nf3, modeord, &w3, &i3);
    for (long k2 = 0; k2 < mt; k2++) {
      long w2 = 0, i2 = 0;
      if (dim > 1) mode_map(k2, mt, nf2, modeord, &w2, &i2);
      for (long k1 = 0; k1 < ms; k1++) {
        long w1, i1;
        mode_map(k1, ms, nf1, modeord, &w1, &i1);
        double kv = h1[i1];
        if (dim > 1) kv *= h2[i2];
        if (dim > 2) kv *= h3[i3];
        const cpx *g = &fk[k1 + k2 * ms + k3 * ms * mt];
        cpx *o = &fw[w1 + w2 * nf1 + w3 * nf1 * nf2];
        o->re = g->re / kv;
        o->im = g->im / kv;
      }
    }
  }
}

/* ------------------------------------------------------------------ FFT (stands in for cuFFT)
 * The reference calls cuFFT (V/include/cufinufft/impl.h:261-297, types.h:108-115), which is not
 * under /root/reference; the published algorithm is the unnormalised DFT
 *   X[k] = sum_n x[n] exp(sign * 2 pi i n k / N).  Restated here as a recursive mixed-radix
 * (2,3,5 + generic) decimation-in-time FFT; nf is always 2^a 3^b 5^c by next235beven. */
static void fft_rec(long n, long stride, const cpx *in, cpx *out, const cpx *tw, long twstride) {
  if (n == 1) { out[0] = in[0]; return; }
  int p = 2;
  if (n % 4 == 0) p = 4; else if (n % 2 == 0) p = 2; else if (n % 3 == 0) p = 3; else if (n % 5 == 0) p = 5;
  else { for (p = 7; n % p; p += 2) {} }
  long m = n / p;
  for (int r = 0; r < p; r++) fft_rec(m, stride * p, in + r * stride, out + r * m, tw, twstride * p);
  /* butterflies: out[k + q m] = sum_r W_n^{r(k+qm)} Y_r[k] */
  cpx tmp[64];
  cpx *t = p <= 64 ? tmp : (cpx *)malloc(sizeof(cpx) * p);
  for (long k = 0; k < m; k++) {
    for (int r = 0; r < p; r++) {
      cpx y = out[r * m + k];
      cpx w = tw[(r * k) * twstride];
      t[r].re = y.re * w.re - y.im * w.im;
      t[r].im = y.re * w.im + y.im * w.re;
    }
    for (int q = 0; q < p; q++) {
      double sr = 0, si = 0;
      for (int r = 0; r < p; r++) {
        /* W_p^{rq} = tw[(r q mod p) * m * twstride] */
        cpx w = tw[((long)(r * q % p)) * m * twstride];
        sr += t[r].re * w.re - t[r].im * w.im;
        si += t[r].re * w.im + t[r].im * w.re;
      }
      out[q * m + k].re = sr;
      out[q * m + k].im = si;
    }
  }
  if (t != tmp) free(t);
}

static void fft_axis(cpx *a, long n, long stride, long nlines_outer, long outer_stride,
                     long nlines_inner, long inner_stride, int sign) {
  cpx *tw = (cpx *)malloc(sizeof(cpx) * (size_t)n);
  for (long k = 0; k < n; k++) {
    double ang = sign * 2.0 * ORC_PI * (double)k / (double)n;
    tw[k].re = cos(ang);
    tw[k].im = sin(ang);
  }
#pragma omp parallel
  {
    cpx *in = (cpx *)malloc(sizeof(cpx) * (size_t)n);
    cpx *out = (cpx *)malloc(sizeof(cpx) * (size_t)n);
#pragma omp for collapse(2) schedule(static)
    for (long o = 0; o < nlines_outer; o++)
      for (long i = 0; i < nlines_inner; i++) {
        cpx *base = a + o * outer_stride + i * inner_stride;
        for (long k = 0; k < n; k++) in[k] = base[k * stride];
        fft_rec(n, 1, in, out, tw, 1);
        for (long k = 0; k < n; k++) base[k * stride] = out[k];
      }
    free(in); free(out);
  }
  free(tw);
}

/* in-place dim-D FFT of fw[(nf3,nf2,nf1)], x fastest; sign=+1/-1 as cufft_ex's direction */
void orc_fft(int dim, long nf1, long nf2, long nf3, cpx *fw, int sign) {
  if (dim < 2) nf2 = 1;
  if (dim < 3) nf3 = 1;
  fft_axis(fw, nf1, 1, nf3, nf1 * nf2, nf2, nf1, sign);
  if (dim > 1) fft_axis(fw, nf2, nf1, nf3, nf1 * nf2, nf1, 1, sign);
  if (dim > 2) fft_axis(fw, nf3, nf1 * nf2, nf2, nf1, nf1, 1, sign);
}

/* ------------------------------------------------------------------ whole transforms */

typedef struct {
  int ns;
  double beta;
  long nf[3];
} orc_info;

/* Types 1 and 2: V/src/cuda/3d/cufinufft3d.cu:18-125 (and 1d/2d twins) under the plan of
 * V/include/cufinufft/impl.h:52-334.  x,y,z may be NULL above `dim`.  n_modes = (ms,mt,mu),
 * ms fastest.  ntransf stacked transforms share the points (c: [ntransf][M], fk:
 * [ntransf][ms*mt*mu]).  Returns the reference's ier (0, 1 = eps too small warning, >1 error). */
int orc_nufft12(int type, int dim, long M, const double *x, const double *y, const double *z,
                cpx *c, int iflag, double eps, const long *n_modes, cpx *fk, int ntransf,
                double upsampfac, int modeord, int kerevalmeth, int prec, orc_info *info) {
  if (type < 1 || type > 2) return 10;
  if (ntransf < 1) return 9;
  if (dim < 1 || dim > 3) return 12;
  int ier;
  double beta;
  if (upsampfac == 0.0) upsampfac = 2.0;
  int ns = orc_setup_spreader(eps, upsampfac, kerevalmeth, prec, &beta, &ier);
  if (ier > 1) return ier;
  long ms = n_modes[0], mt = dim > 1 ? n_modes[1] : 1, mu = dim > 2 ? n_modes[2] : 1;
  long nf1 = orc_set_nf_type12(ms, upsampfac, ns);
  long nf2 = dim > 1 ? orc_set_nf_type12(mt, upsampfac, ns) : 1;
  long nf3 = dim > 2 ? orc_set_nf_type12(mu, upsampfac, ns) : 1;
  if (info) { info->ns = ns; info->beta = beta; info->nf[0] = nf1; info->nf[1] = nf2; info->nf[2] = nf3; }
  double *h1 = (double *)malloc(sizeof(double) * (size_t)(nf1 / 2 + 1));
  double *h2 = (double *)malloc(sizeof(double) * (size_t)(nf2 / 2 + 1));
  double *h3 = (double *)malloc(sizeof(double) * (size_t)(nf3 / 2 + 1));
  orc_fseries(nf1, ns, beta, h1);
  if (dim > 1) orc_fseries(nf2, ns, beta, h2);
  if (dim > 2) orc_fseries(nf3, ns, beta, h3);
  long nf = nf1 * nf2 * nf3, N = ms * mt * mu;
  cpx *fw = (cpx *)malloc(sizeof(cpx) * (size_t)nf);
  int sign = iflag >= 0 ? 1 : -1;
  for (int t = 0; t < ntransf; t++) {
    if (type == 1) {
      memset(fw, 0, sizeof(cpx) * (size_t)nf);
      orc_spread(dim, M, x, y, z, c + (size_t)t * M, nf1, nf2, nf3, ns, beta, prec, fw);
      orc_fft(dim, nf1, nf2, nf3, fw, sign);
      orc_deconvolve(dim, ms, mt, mu, nf1, nf2, nf3, fw, fk + (size_t)t * N, h1, h2, h3, modeord);
    } else {
      orc_amplify(dim, ms, mt, mu, nf1, nf2, nf3, fw, fk + (size_t)t * N, h1, h2, h3, modeord);
      orc_fft(dim, nf1, nf2, nf3, fw, sign);
      orc_interp(dim, M, x, y, z, c + (size_t)t * M, nf1, nf2, nf3, ns, beta, prec, fw);
    }
  }
  free(fw); free(h1); free(h2); free(h3);
  return ier;
}

/* V/include/cufinufft/utils.h:126-152 (arrayrange/arraywidcen); GROWFRAC=0.1
 * (V/include/finufft_common/constants.h) */
static void arraywidcen(long n, const double *a, double *w, double *c) {
  double lo = INFINITY, hi = -INFINITY;
  for (long i = 0; i < n; i++) { if (a[i] < lo) lo = a[i]; if (a[i] > hi) hi = a[i]; }
  *w = (hi - lo) / 2;
  *c = (hi + lo) / 2;
  if (fabs(*c) < 0.1 * (*w)) { *w += fabs(*c); *c = 0.0; }
}

/* V/include/cufinufft/utils.h:154-183 (set_nhg_type3) */
static void set_nhg_type3(double S, double X, double upsampfac, int ns, long *nf, double *h,
                          double *gam) {
  int nss = ns + 1;
  double Xsafe = X, Ssafe = S;
  if (X == 0.0) {
    if (S == 0.0) { Xsafe = 1.0; Ssafe = 1.0; }
    else Xsafe = fmax(Xsafe, 1.0 / S);
  } else Ssafe = fmax(Ssafe, 1.0 / X);
  double nfd = 2.0 * upsampfac * Ssafe * Xsafe / ORC_PI + nss;
  if (!isfinite(nfd)) nfd = 0.0;
  long n = (long)(int)nfd;
  if (n < 2 * ns) n = 2 * ns;
  n = orc_next235beven(n, 1);
  *nf = n;
  *h = 2 * ORC_PI / (double)n;
  *gam = (double)n / (2.0 * upsampfac * Ssafe);
}

/* Type 3: V/include/cufinufft/impl.h:461-823 (setpts) + V/src/cuda/3d/cufinufft3d.cu:127-183
 * (exec); the CPU statement of the same maths is V/src/finufft_core.cpp:1047-1203.
 * x,y,z: M sources; s,t,u: N targets; c: [ntransf][M]; f: [ntransf][N]. */
int orc_nufft3(int dim, long M, const double *x, const double *y, const double *z, const cpx *c,
               int iflag, double eps, long N, const double *s, const double *t, const double *u,
               cpx *f, int ntransf, double upsampfac, int kerevalmeth, int prec, orc_info *info) {
  if (dim < 1 || dim > 3) return 12;
  if (ntransf < 1) return 9;
  int ier;
  double beta;
  if (upsampfac == 0.0) upsampfac = (eps >= 1e-9) ? 1.25 : 2.0; /* impl.h:151-155, type 3 */
  int ns = orc_setup_spreader(eps, upsampfac, kerevalmeth, prec, &beta, &ier);
  if (ier > 1) return ier;
  const double *X[3] = {x, y, z}, *S[3] = {s, t, u};
  double Xw[3] = {0, 0, 0}, C[3] = {0, 0, 0}, Sw[3] = {0, 0, 0}, D[3] = {0, 0, 0}, h[3] = {1, 1, 1}, gam[3] = {1, 1, 1};
  long nf[3] = {1, 1, 1};
  for (int d = 0; d < dim; d++) {
    arraywidcen(M, X[d], &Xw[d], &C[d]);
    arraywidcen(N, S[d], &Sw[d], &D[d]);
    set_nhg_type3(Sw[d], Xw[d], upsampfac, ns, &nf[d], &h[d], &gam[d]);
  }
  if (info) { info->ns = ns; info->beta = beta; info->nf[0] = nf[0]; info->nf[1] = nf[1]; info->nf[2] = nf[2]; }
  /* rescaled sources x' = (x-C)/gam, rescaled targets s' = h gam (s-D)  (impl.h:609-686) */
  double *xp[3] = {0, 0, 0}, *sp[3] = {0, 0, 0};
  for (int d = 0; d < dim; d++) {
    xp[d] = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1));
    sp[d] = (double *)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
    double ig = 1.0 / gam[d];
    for (long j = 0; j < M; j++) xp[d][j] = (X[d][j] - C[d]) * ig;
    double sc = h[d] * gam[d];
    for (long k = 0; k < N; k++) sp[d][k] = sc * (S[d][k] - D[d]);
  }
  double sgn = iflag >= 0 ? 1.0 : -1.0;
  /* prephase_j = cis(sgn * D.x_j)   (impl.h:632-663) */
  cpx *pre = (cpx *)malloc(sizeof(cpx) * (size_t)(M > 0 ? M : 1));
  int anyD = D[0] != 0 || D[1] != 0 || D[2] != 0;
  for (long j = 0; j < M; j++) {
    if (anyD) {
      double ph = 0;
      for (int d = 0; d < dim; d++) ph += D[d] * X[d][j];
      pre[j].re = cos(ph);
      pre[j].im = sgn * sin(ph);
    } else { pre[j].re = 1; pre[j].im = 0; }
  }
  /* deconv_k = cis(sgn * C.(s_k-D)) / prod phihat_d(s'_k)   (impl.h:688-773) */
  cpx *dec = (cpx *)malloc(sizeof(cpx) * (size_t)(N > 0 ? N : 1));
  double *ph1 = (double *)malloc(sizeof(double) * (size_t)(N > 0 ? N : 1));
  for (long k = 0; k < N; k++) { dec[k].re = 1.0; dec[k].im = 0.0; }
  for (int d = 0; d < dim; d++) {
    orc_nuft(ns, beta, N, sp[d], ph1);
    for (long k = 0; k < N; k++) dec[k].re *= ph1[k];
  }
  int anyC = (C[0] != 0 || C[1] != 0 || C[2] != 0) && isfinite(C[0]) && isfinite(C[1]) && isfinite(C[2]);
  for (long k = 0; k < N; k++) {
    double inv = 1.0 / dec[k].re;
    if (anyC) {
      double ph = 0;
      for (int d = 0; d < dim; d++) ph += C[d] * (S[d][k] - D[d]);
      dec[k].re = cos(ph) * inv;
      dec[k].im = sgn * sin(ph) * inv;
    } else { dec[k].re = inv; dec[k].im = 0; }
  }
  free(ph1);
  long nftot = nf[0] * nf[1] * nf[2];
  cpx *fw = (cpx *)malloc(sizeof(cpx) * (size_t)nftot);
  cpx *cp = (cpx *)malloc(sizeof(cpx) * (size_t)(M > 0 ? M : 1));
  int rc = ier;
  for (int tr = 0; tr < ntransf; tr++) {
    const cpx *ct = c + (size_t)tr * M;
    for (long j = 0; j < M; j++) {
      cp[j].re = ct[j].re * pre[j].re - ct[j].im * pre[j].im;
      cp[j].im = ct[j].re * pre[j].im + ct[j].im * pre[j].re;
    }
    memset(fw, 0, sizeof(cpx) * (size_t)nftot);
    orc_spread(dim, M, xp[0], xp[1], xp[2], cp, nf[0], nf[1], nf[2], ns, beta, prec, fw);
    /* inner type 2 with modes (nf1,nf2,nf3), modeord 0, same iflag/eps/sigma (impl.h:795-812) */
    cpx *ft = f + (size_t)tr * N;
    int r2 = orc_nufft12(2, dim, N, sp[0], sp[1], sp[2], ft, iflag, eps, nf, fw, 1, upsampfac, 0,
                         kerevalmeth, prec, NULL);
    if (r2 > 1) { rc = r2; break; }
    for (long k = 0; k < N; k++) {
      double a = ft[k].re, b = ft[k].im;
      ft[k].re = a * dec[k].re - b * dec[k].im;
      ft[k].im = a * dec[k].im + b * dec[k].re;
    }
  }
  for (int d = 0; d < dim; d++) { free(xp[d]); free(sp[d]); }
  free(pre); free(dec); free(fw); free(cp);
  return rc;
}

/* ------------------------------------------------------------------ direct NUDFT (ground truth)
 * Definitions: V/docs/math.rst:41-58,84-87; loops as tests/ops_test.py:43-45,78-84,113-116 and
 * V/test/utils/dirft3d.hpp.  modeord as get_frequency_array (src/jax_finufft/ops.py:114-123). */
static inline long freq_of(long k, long n, int modeord) {
  if (modeord == 0) return k - n / 2;
  return k < (n + 1) / 2 ? k : k - n;
}

void orc_dirft1(int dim, long M, const double *x, const double *y, const double *z, const cpx *c,
                int iflag, long ms, long mt, long mu, cpx *fk, int modeord) {
  if (dim < 2) mt = 1;
  if (dim < 3) mu = 1;
  double sg = iflag >= 0 ? 1.0 : -1.0;
  long N = ms * mt * mu;
#pragma omp parallel for schedule(static)
  for (long n = 0; n < N; n++) {
    long k1 = freq_of(n % ms, ms, modeord);
    long k2 = dim > 1 ? freq_of((n / ms) % mt, mt, modeord) : 0;
    long k3 = dim > 2 ? freq_of(n / (ms * mt), mu, modeord) : 0;
    double sr = 0, si = 0;
    for (long j = 0; j < M; j++) {
      double ph = (double)k1 * x[j];
      if (dim > 1) ph += (double)k2 * y[j];
      if (dim > 2) ph += (double)k3 * z[j];
      double cr = cos(ph), ci = sg * sin(ph);
      sr += c[j].re * cr - c[j].im * ci;
      si += c[j].re * ci + c[j].im * cr;
    }
    fk[n].re = sr;
    fk[n].im = si;
  }
}

void orc_dirft2(int dim, long M, const double *x, const double *y, const double *z, cpx *c,
                int iflag, long ms, long mt, long mu, const cpx *fk, int modeord) {
  if (dim < 2) mt = 1;
  if (dim < 3) mu = 1;
  double sg = iflag >= 0 ? 1.0 : -1.0;
  long N = ms * mt * mu;
#pragma omp parallel for schedule(static)
  for (long j = 0; j < M; j++) {
    double sr = 0, si = 0;
    for (long n = 0; n < N; n++) {
      long k1 = freq_of(n % ms, ms, modeord);
      long k2 = dim > 1 ? freq_of((n / ms) % mt, mt, modeord) : 0;
      long k3 = dim > 2 ? freq_of(n / (ms * mt), mu, modeord) : 0;
      double ph = (double)k1 * x[j];
      if (dim > 1) ph += (double)k2 * y[j];
      if (dim > 2) ph += (double)k3 * z[j];
      double cr = cos(ph), ci = sg * sin(ph);
      sr += fk[n].re * cr - fk[n].im * ci;
      si += fk[n].re * ci + fk[n].im * cr;
    }
    c[j].re = sr;
    c[j].im = si;
  }
}

void orc_dirft3(int dim, long M, const double *x, const double *y, const double *z, const cpx *c,
                int iflag, long N, const double *s, const double *t, const double *u, cpx *f) {
  double sg = iflag >= 0 ? 1.0 : -1.0;
#pragma omp parallel for schedule(static)
  for (long k = 0; k < N; k++) {
    double sr = 0, si = 0;
    for (long j = 0; j < M; j++) {
      double ph = s[k] * x[j];
      if (dim > 1) ph += t[k] * y[j];
      if (dim > 2) ph += u[k] * z[j];
      double cr = cos(ph), ci = sg * sin(ph);
      sr += c[j].re * cr - c[j].im * ci;
      si += c[j].re * ci + c[j].im * cr;
    }
    f[k].re = sr;
    f[k].im = si;
  }
}

/* torchrun exports OMP_NUM_THREADS=1 to its workers; the CPU baseline uses every core it is given */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
