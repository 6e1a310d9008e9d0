// stages.cu -- the streaming (HBM-bound) stages around spread/interp/FFT:
//   * kernel Fourier-series evaluation           (V/src/cuda/common.cu:27-60,104-125)
//   * type-1 step 3: deconvolve + crop           (V/src/cuda/deconvolve_wrapper.cu:16-118,291-315)
//   * type-2 step 1: amplify + zero-pad, ONE pass (V/src/cuda/deconvolve_wrapper.cu:121-223,316-325
//                                                  -- the reference memsets all of fw, then scatters)
//   * type-3 setpts: min/max, rescale, prephase, kernel FT at targets, deconv factors
//                                                 (V/include/cufinufft/impl.h:509-773,
//                                                  V/src/cuda/common.cu:68-102, utils.h:126-152)
// Algorithmic bytes: deconvolve 16 B/mode; amplify 8 B/mode + 8 B/fine cell.
#include "plan.h"

namespace b2n {

// ------------------------------------------------------------------ kernel Fourier series
struct QuadArgs {
  double f[3][32];
  double ph[3][32];
  int q;
  int nf[3];
};

template <typename T>
__global__ void k_fseries(const QuadArgs qa, T *o1, T *o2, T *o3) {
  const int d = blockIdx.y;
  T *out = d == 0 ? o1 : (d == 1 ? o2 : o3);
  const int n = qa.nf[d] / 2 + 1;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < qa.q; k++) s += qa.f[d][k] * 2.0 * cos((double)i * qa.ph[d][k]);
    out[i] = (T)((i & 1) ? -s : s);
  }
}

template <typename T> int compute_fseries(Plan<T> &p) {
  QuadArgs qa;
  int nmax = 0;
  for (int d = 0; d < p.dim; d++) {
    qa.q = kernel_quadrature(p.ns, p.beta, false, p.nf[d], qa.f[d], qa.ph[d]);
    qa.nf[d] = (int)p.nf[d];
    nmax = std::max(nmax, (int)p.nf[d] / 2 + 1);
  }
  dim3 grid(std::min(cdiv(nmax, 128), 1024), p.dim);
  k_fseries<T><<<grid, 128, 0, p.stream>>>(qa, p.fwker[0], p.fwker[1], p.fwker[2]);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ mode <-> fine-grid maps
struct ModeGeom {
  int dim, modeord;
  int ms[3];
  int nf[3];
};

// output mode index k (0..ms-1) -> fine index w and |frequency| (deconvolve_wrapper.cu:91-111)
__device__ __forceinline__ void mode_to_fine(int k, int ms, int nf, int modeord, int &w, int &a) {
  int p;
  if (modeord == 0) p = k - ms / 2;
  else p = (k >= ms - ms / 2) ? k - ms : k;
  w = p >= 0 ? p : nf + p;
  a = p >= 0 ? p : -p;
}
// fine index w -> mode index k or -1 if that cell carries no mode
__device__ __forceinline__ int fine_to_mode(int w, int ms, int nf, int modeord, int &a) {
  int p;
  if (w <= (ms - 1) / 2) p = w;
  else if (w >= nf - ms / 2) p = w - nf;
  else return -1;
  a = p >= 0 ? p : -p;
  if (modeord == 0) return p + ms / 2;
  return p >= 0 ? p : ms + p;
}

// A CTA owns a tile of `cols` modes x 256/cols output rows (cols = a power of two <= 256 covering
// ms[0], so short rows still fill the CTA); the row decode uses 32-bit arithmetic (the first
// version decoded every mode with 64-bit divisions).
template <typename T>
__global__ void __launch_bounds__(256) k_deconvolve(const ModeGeom g, const cpx<T> *__restrict__ fw,
                                                     cpx<T> *__restrict__ fk,
                                                     const T *__restrict__ h1,
                                                     const T *__restrict__ h2,
                                                     const T *__restrict__ h3, int64_t nmodes,
                                                     int64_t nftot, int nxchunk, int lgcols, unsigned rows) {
  const int cols = 1 << lgcols, rpb = 256 >> lgcols;
  const unsigned rb = blockIdx.x / (unsigned)nxchunk;
  const int xc = (int)(blockIdx.x - rb * (unsigned)nxchunk);
  const unsigned row = rb * rpb + (threadIdx.x >> lgcols);
  const int k1 = xc * cols + (threadIdx.x & (cols - 1));
  if (k1 >= g.ms[0] || row >= rows) return;
  fw += (int64_t)blockIdx.y * nftot;
  fk += (int64_t)blockIdx.y * nmodes + (int64_t)row * g.ms[0];
  int w1, a1, w2 = 0, a2 = 0, w3 = 0, a3 = 0;
  mode_to_fine(k1, g.ms[0], g.nf[0], g.modeord, w1, a1);
  T kv = h1[a1];  // same association as before: ((h1 * h2) * h3)
  if (g.dim > 1) {
    const unsigned k3 = row / (unsigned)g.ms[1], k2 = row - k3 * (unsigned)g.ms[1];
    mode_to_fine((int)k2, g.ms[1], g.nf[1], g.modeord, w2, a2);
    kv *= h2[a2];
    if (g.dim > 2) {
      mode_to_fine((int)k3, g.ms[2], g.nf[2], g.modeord, w3, a3);
      kv *= h3[a3];
    }
  }
  const cpx<T> v = fw[((int64_t)w3 * g.nf[1] + w2) * g.nf[0] + w1];
  cpx<T> o;
  o.x = v.x / kv;
  o.y = v.y / kv;
  fk[k1] = o;
}

template <typename T> int deconvolve(Plan<T> &p, const cpx<T> *fw, cpx<T> *fk, int ntr) {
  ModeGeom g;
  g.dim = p.dim;
  g.modeord = p.opts.modeord;
  for (int d = 0; d < 3; d++) { g.ms[d] = (int)p.ms[d]; g.nf[d] = (int)p.nf[d]; }
  int lgcols = 0;
  while (lgcols < 8 && (1 << lgcols) < p.ms[0]) lgcols++;
  const int cols = 1 << lgcols, rpb = 256 >> lgcols;
  const int nxchunk = cdiv(p.ms[0], cols);
  const int64_t rows = p.nmodes / p.ms[0];
  const int64_t nblk = (int64_t)cdiv(rows, rpb) * nxchunk;
  if (nblk > 0x7fffffffLL || rows > 0x7fffffffLL) return B2N_ERR_NDATA_NOTVALID;
  dim3 grid((unsigned)nblk, (unsigned)ntr);
  k_deconvolve<T><<<grid, 256, 0, p.stream>>>(g, fw, fk, p.fwker[0], p.fwker[1], p.fwker[2],
                                              p.nmodes, p.nftot, nxchunk, lgcols, (unsigned)rows);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

// One pass over the fine grid: every cell gets either its amplified mode or zero.  A CTA owns a
// 512-cell piece of one fine-grid row: the row's (y, z) decode and mode test happen once per CTA
// with 32-bit arithmetic, three rows out of four (3-D) are pure zero fill, and every thread
// stores two adjacent cells as one 16-byte word.  (The first version decoded each cell with
// 64-bit divisions: 0.60 ms for 1.15 GB at C3, issue-bound at 29 % of the HBM roofline.)
constexpr int AMP_T = 256, AMP_TMAX = 512;
template <typename T>
__global__ void __launch_bounds__(AMP_TMAX) k_amplify(const ModeGeom g, cpx<T> *__restrict__ fw,
                                                    const cpx<T> *__restrict__ fk,
                                                    const T *__restrict__ h1, const T *__restrict__ h2,
                                                    const T *__restrict__ h3, int64_t nmodes,
                                                    int64_t nftot, int nxchunk) {
  const unsigned row = blockIdx.x / (unsigned)nxchunk;
  const int xc = (int)(blockIdx.x - row * (unsigned)nxchunk);
  fw += (int64_t)blockIdx.y * nftot + (int64_t)row * g.nf[0];
  fk += (int64_t)blockIdx.y * nmodes;
  int a2 = 0, a3 = 0, k2 = 0, k3 = 0;
  bool rowok = true;
  if (g.dim > 1) {
    const unsigned w3 = row / (unsigned)g.nf[1], w2 = row - w3 * (unsigned)g.nf[1];
    k2 = fine_to_mode((int)w2, g.ms[1], g.nf[1], g.modeord, a2);
    rowok = k2 >= 0;
    if (rowok && g.dim > 2) {
      k3 = fine_to_mode((int)w3, g.ms[2], g.nf[2], g.modeord, a3);
      rowok = k3 >= 0;
    }
  }
  const int w1 = xc * (2 * (int)blockDim.x) + 2 * threadIdx.x;  // nf[0] is even for every grid the library builds
  if (w1 >= g.nf[0]) return;
  cpx<T> o[2];
  o[0].x = o[0].y = o[1].x = o[1].y = T(0);
  if (rowok) {
    const T hv2 = g.dim > 1 ? h2[a2] : T(1), hv3 = g.dim > 2 ? h3[a3] : T(1);
    const cpx<T> *fkrow = fk + ((int64_t)k3 * g.ms[1] + k2) * g.ms[0];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      int a1;
      const int k1 = fine_to_mode(w1 + e, g.ms[0], g.nf[0], g.modeord, a1);
      if (k1 >= 0) {
        T kv = h1[a1];  // same association as the reference: ((h1 * h2) * h3)
        if (g.dim > 1) kv *= hv2;
        if (g.dim > 2) kv *= hv3;
        const cpx<T> v = fkrow[k1];
        o[e].x = v.x / kv;
        o[e].y = v.y / kv;
      }
    }
  }
  if (sizeof(T) == 4 && (g.nf[0] & 1) == 0) {
    *reinterpret_cast<float4 *>(fw + w1) = make_float4((float)o[0].x, (float)o[0].y, (float)o[1].x, (float)o[1].y);
  } else {
    fw[w1] = o[0];
    if (w1 + 1 < g.nf[0]) fw[w1 + 1] = o[1];
  }
}

template <typename T> int amplify(Plan<T> &p, cpx<T> *fw, const cpx<T> *fk, int ntr) {
  ModeGeom g;
  g.dim = p.dim;
  g.modeord = p.opts.modeord;
  for (int d = 0; d < 3; d++) { g.ms[d] = (int)p.ms[d]; g.nf[d] = (int)p.nf[d]; }
  // one thread per cell pair; rows that are a multiple of 512 cells are cut into 512-cell pieces,
  // any other row into equal pieces of up to 512 pairs: with fixed 512-cell pieces a 540-cell row
  // (type 3's inner grid at C5) took two CTAs, the second one with 14 live threads -- 0.65 ms for
  // 1.26 GB
  const int pairs = (int)((p.nf[0] + 1) / 2);
  const bool even_pieces = pairs % AMP_T == 0;
  const int nxchunk = even_pieces ? pairs / AMP_T : cdiv(pairs, AMP_TMAX);
  const int threads = even_pieces ? AMP_T : std::max(64, (cdiv(pairs, nxchunk) + 31) / 32 * 32);
  const int64_t rows = p.nftot / p.nf[0];
  if (rows * nxchunk > 0x7fffffffLL) return B2N_ERR_NDATA_NOTVALID;
  dim3 grid((unsigned)(rows * nxchunk), (unsigned)ntr);
  k_amplify<T><<<grid, threads, 0, p.stream>>>(g, fw, fk, p.fwker[0], p.fwker[1], p.fwker[2], p.nmodes,
                                             p.nftot, nxchunk);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

// ------------------------------------------------------------------ type 3: min/max of 6 arrays
template <typename T> struct MinMaxArgs {
  const T *a[6];
  int64_t n[6];
};

template <typename T>
__global__ void __launch_bounds__(256) k_minmax(const MinMaxArgs<T> mm, double *part /*[6][grid][2]*/) {
  const int arr = blockIdx.y;
  const T *a = mm.a[arr];
  const int64_t n = mm.n[arr];
  T lo = INFINITY, hi = -INFINITY;
  if (a) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
         i += (int64_t)gridDim.x * blockDim.x) {
      const T v = a[i];
      lo = v < lo ? v : lo;
      hi = v > hi ? v : hi;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const T l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  __shared__ T slo[8], shi[8];
  if ((threadIdx.x & 31) == 0) { slo[threadIdx.x >> 5] = lo; shi[threadIdx.x >> 5] = hi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; k++) {
      lo = slo[k] < lo ? slo[k] : lo;
      hi = shi[k] > hi ? shi[k] : hi;
    }
    part[((int64_t)arr * gridDim.x + blockIdx.x) * 2 + 0] = (double)lo;
    part[((int64_t)arr * gridDim.x + blockIdx.x) * 2 + 1] = (double)hi;
  }
}

// ONE fused reduction + ONE readback for all 2*dim arrays (the reference runs
// thrust::minmax_element and two D2H copies per array: utils.h:126-135).
template <typename T>
int t3_minmax(cudaStream_t st, int dim, int64_t M, const T *const *x, int64_t N, const T *const *s,
              double *lohi) {
  MinMaxArgs<T> mm;
  for (int d = 0; d < 3; d++) {
    mm.a[d] = d < dim ? x[d] : nullptr;
    mm.n[d] = M;
    mm.a[3 + d] = d < dim ? s[d] : nullptr;
    mm.n[3 + d] = N;
  }
  const int nb = 148 * 2;
  double *part = nullptr;
  if (int e = dev_alloc_t(&part, (size_t)6 * nb * 2, st)) return e;
  k_minmax<T><<<dim3(nb, 6), 256, 0, st>>>(mm, part);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  std::vector<double> h((size_t)6 * nb * 2);
  B2N_CUDA_OK(cudaMemcpyAsync(h.data(), part, h.size() * sizeof(double), cudaMemcpyDeviceToHost, st));
  B2N_CUDA_OK(cudaStreamSynchronize(st));
  dev_free(part, st);
  for (int a = 0; a < 6; a++) {
    double lo = INFINITY, hi = -INFINITY;
    for (int b = 0; b < nb; b++) {
      lo = std::min(lo, h[((size_t)a * nb + b) * 2]);
      hi = std::max(hi, h[((size_t)a * nb + b) * 2 + 1]);
    }
    lohi[2 * a] = lo;
    lohi[2 * a + 1] = hi;
  }
  return 0;
}

// ------------------------------------------------------------------ type 3: fused setpts kernels
template <typename T> struct T3Args {
  const T *x[3];
  T *xp[3];
  const T *s[3];
  T *sp[3];
  cpx<T> *prephase, *deconv;
  T C[3], ig[3], D[3], sc[3];  // centre, 1/gamma, target centre, h*gamma
  T f[3][32], z[3][32];        // kernel-FT quadrature (same for every dim, kept per dim for clarity)
  int q, dim;
  int anyC, anyD;
  T sign;
  int64_t M, N;
};

// sources: x' = (x - C)/gamma; prephase = cis(sign * D.x)           (impl.h:609-663)
template <typename T> __global__ void __launch_bounds__(256) k_t3_sources(const T3Args<T> a) {
  const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= a.M) return;
  T ph = T(0);
  for (int d = 0; d < a.dim; d++) {
    const T x = a.x[d][j];
    a.xp[d][j] = (x + (-a.C[d])) * a.ig[d];
    ph += a.D[d] * x;
  }
  if (a.anyD) {
    cpx<T> v;
    v.x = cos(ph);
    v.y = sin(ph) * a.sign;
    a.prephase[j] = v;
  }
}

// targets: s' = h gamma (s - D); deconv = cis(sign * C.(s-D)) / prod_d phihat(s'_d)
//                                                                    (impl.h:665-773, common.cu:68-102)
template <typename T> __global__ void __launch_bounds__(256) k_t3_targets(const T3Args<T> a) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= a.N) return;
  T prod = T(1), ph = T(0);
  for (int d = 0; d < a.dim; d++) {
    const T s = a.s[d][k];
    const T spv = a.sc[d] * (s + (-a.D[d]));
    a.sp[d][k] = spv;
    T acc = T(0);
    for (int n = 0; n < a.q; n++) acc += a.f[d][n] * T(2) * cos(spv * a.z[d][n]);
    prod *= acc;
    ph += a.C[d] * (s + (-a.D[d]));
  }
  cpx<T> v;
  v.x = T(1) / prod;
  v.y = T(0);
  if (a.anyC) {
    const T inv = v.x;
    v.x = cos(ph) * inv;
    v.y = a.sign * sin(ph) * inv;
  }
  a.deconv[k] = v;
}

template <typename T> int t3_prepare(Plan<T> &p, const T *const *x, const T *const *s) {
  T3Args<T> a;
  double f[MAX_NQUAD], z[MAX_NQUAD];
  a.q = kernel_quadrature(p.ns, p.beta, true, 0, f, z);
  a.dim = p.dim;
  a.M = p.pts.M;
  a.N = p.N3;
  a.anyC = 0;
  a.anyD = 0;
  bool finiteC = true;
  for (int d = 0; d < 3; d++) {
    a.x[d] = d < p.dim ? x[d] : nullptr;
    a.s[d] = d < p.dim ? s[d] : nullptr;
    a.xp[d] = p.xp[d];
    a.sp[d] = p.sp[d];
    a.C[d] = (T)p.t3C[d];
    a.D[d] = (T)p.t3D[d];
    a.ig[d] = T(1) / (T)p.t3gam[d];
    a.sc[d] = (T)p.t3h[d] * (T)p.t3gam[d];
    if (p.t3C[d] != 0) a.anyC = 1;
    if (p.t3D[d] != 0) a.anyD = 1;
    if (!std::isfinite(p.t3C[d])) finiteC = false;
    for (int n = 0; n < a.q; n++) {
      // the reference rounds z to T first, then evaluates the kernel there (common.cu:218-221)
      const T zT = (T)z[n];
      a.z[d][n] = zT;
      a.f[d][n] = (T)f[n];
    }
  }
  if (!finiteC) a.anyC = 0;
  a.prephase = p.prephase;
  a.deconv = p.deconv;
  a.sign = p.iflag >= 0 ? T(1) : T(-1);
  if (a.M > 0) k_t3_sources<T><<<cdiv(a.M, 256), 256, 0, p.stream>>>(a);  B2N_LAUNCHED(1);
  if (a.N > 0) k_t3_targets<T><<<cdiv(a.N, 256), 256, 0, p.stream>>>(a);  B2N_LAUNCHED(1);
  B2N_LAUNCH_OK();
  return 0;
}

template int compute_fseries<float>(Plan<float> &);
template int compute_fseries<double>(Plan<double> &);
template int deconvolve<float>(Plan<float> &, const cpx<float> *, cpx<float> *, int);
template int deconvolve<double>(Plan<double> &, const cpx<double> *, cpx<double> *, int);
template int amplify<float>(Plan<float> &, cpx<float> *, const cpx<float> *, int);
template int amplify<double>(Plan<double> &, cpx<double> *, const cpx<double> *, int);
template int t3_minmax<float>(cudaStream_t, int, int64_t, const float *const *, int64_t, const float *const *, double *);
template int t3_minmax<double>(cudaStream_t, int, int64_t, const double *const *, int64_t, const double *const *, double *);
template int t3_prepare<float>(Plan<float> &, const float *const *, const float *const *);
template int t3_prepare<double>(Plan<double> &, const double *const *, const double *const *);

}  // namespace b2n
