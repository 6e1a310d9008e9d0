// api.cu -- plan lifecycle, the type 1/2/3 pipelines, the per-process plan cache and the C ABI.
// Replaces lib/kernels.cc.cu (run_nufft), lib/cufinufft_wrapper.* and cuFINUFFT's
// makeplan/setpts/execute/destroy (V/include/cufinufft/impl.h, V/src/cuda/{1,2,3}d/cufinufft*d.cu).
#include <atomic>
#include <algorithm>
#include <cmath>
#include <cstring>
#include <mutex>
#include <new>

#include "plan.h"

namespace b2n {

std::atomic<unsigned long long> g_launch_count{0};

// setpts cache switch (sort.cu: binsort_points): -1 = not decided yet, read B2N_SETPTS_CACHE once
static std::atomic<int> g_setpts_cache{-1};
bool setpts_cache_enabled() {
  int v = g_setpts_cache.load(std::memory_order_relaxed);
  if (v < 0) {
    const char *e = getenv("B2N_SETPTS_CACHE");
    v = (e && e[0] && e[0] != '0') ? 1 : 0;
    g_setpts_cache.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

// ------------------------------------------------------------------------------------ utilities
struct StageTimer {
  cudaEvent_t a = nullptr, b = nullptr;
  cudaStream_t st;
  bool on;
  double *slot;
  StageTimer(bool enable, cudaStream_t s, double *dst) : st(s), on(enable), slot(dst) {
    if (on) {
      cudaEventCreate(&a);
      cudaEventCreate(&b);
      cudaEventRecord(a, st);
    }
  }
  ~StageTimer() {
    if (on) {
      cudaEventRecord(b, st);
      cudaEventSynchronize(b);
      float ms = 0;
      cudaEventElapsedTime(&ms, a, b);
      *slot += ms;
      cudaEventDestroy(a);
      cudaEventDestroy(b);
    }
  }
};

template <typename T> static size_t tile_bytes_for(int dim, int ns, const int *bin) {
  return tile_smem_bytes<T>(dim, ns, bin);
}

static const int CAND2[][3] = {{32, 32, 1}, {32, 16, 1}, {16, 16, 1}, {16, 8, 1}, {8, 8, 1}, {8, 4, 1}, {4, 4, 1}, {2, 2, 1}};
static const int CAND3[][3] = {{16, 16, 8}, {16, 8, 8}, {8, 8, 8}, {8, 8, 4}, {8, 4, 4}, {4, 4, 4}, {4, 4, 2}, {2, 2, 2}};
constexpr size_t SMEM_3CTA = 76000, SMEM_2CTA = 115000, SMEM_1CTA = 232448;

// B200 bin-size choice (replaces cufinufft_setup_binsize, V/src/cuda/common.cu:522-641): the
// largest candidate whose tile + weight batch lets 3 CTAs share an SM's 227 KB; failing that 2,
// then 1.  Returns false if no tile fits (caller falls back to the GM kernels).
template <typename T> static bool choose_bins(int dim, int ns, int *bin) {
  if (dim == 1) {
    bin[0] = 1024; bin[1] = 1; bin[2] = 1;
    return true;
  }
  const int(*cand)[3] = dim == 2 ? CAND2 : CAND3;
  const size_t limits[3] = {SMEM_3CTA, SMEM_2CTA, SMEM_1CTA};
  for (size_t lim : limits)
    for (int i = 0; i < 8; i++) {
      size_t b = tile_bytes_for<T>(dim, ns, cand[i]);
      if (b > 0 && b <= lim) {
        bin[0] = cand[i][0]; bin[1] = cand[i][1]; bin[2] = cand[i][2];
        return true;
      }
    }
  return false;
}

template <typename T> static void fill_table(HornerTable<T> &tab, int ns, double beta, bool direct, int *ncoef) {
  double coef[MAX_NCOEF_H * 16];
  int nc = horner_fit(ns, beta, sizeof(T) == 8, coef);
  std::memset(&tab, 0, sizeof(tab));
  for (int k = 0; k < nc; k++)
    for (int j = 0; j < ns; j++) tab.c[k][j] = (T)coef[k * 16 + j];
  tab.ncoef = nc;
  tab.ns = ns;
  tab.es_c = (T)(4.0 / ((double)ns * ns));
  tab.es_beta = (T)beta;
  tab.direct = direct ? 1 : 0;
  *ncoef = nc;
}

// ------------------------------------------------------------------------------------ Plan
template <typename T> Plan<T>::~Plan() {
  cudaStream_t st = stream;
  for (int d = 0; d < 3; d++) {
    dev_free(fwker[d], st);
    dev_free(xp[d], st);
    dev_free(sp[d], st);
  }
  dev_free(fw, st);
  dev_free(cpack, st);
  dev_free(fw_stack, st);
  if (pts.ovf_host) cudaFreeHost(pts.ovf_host);
  dev_free(pts.rec, st);
  dev_free(pts.tmp, st);
  dev_free(pts.idx, st);
  dev_free(pts.key_cnt, st);
  dev_free(pts.key_start, st);
  dev_free(pts.bucket_cur, st);
  dev_free(pts.bin_start, st);
  dev_free(pts.sp_off, st);
  dev_free(pts.sp_bin, st);
  dev_free(pts.sig, st);
  dev_free(prephase, st);
  dev_free(deconv, st);
  if (side) cudaStreamDestroy(side);
  if (ev_fork) cudaEventDestroy(ev_fork);
  if (ev_join) cudaEventDestroy(ev_join);
  if (has_fft) cufftDestroy(fft);
  if (pruned) { cufftDestroy(fft_z); for (int k = 0; k < 2; k++) if (slab_n[k]) cufftDestroy(fft_xy[k]); }
  delete inner;
}

template <typename T> void Plan<T>::set_stream(cudaStream_t s) {
  stream = s;
  if (has_fft) cufftSetStream(fft, s);
  if (pruned) { cufftSetStream(fft_z, s); for (int k = 0; k < 2; k++) if (slab_n[k]) cufftSetStream(fft_xy[k], s); }
  if (inner) inner->set_stream(s);
}

template <typename T>
int Plan<T>::init(int type_, int dim_, const int64_t *n_modes, int iflag_, int ntransf_,
                  double eps_, const b2n_opts *o) {
  is_double = sizeof(T) == 8;
  if (type_ < 1 || type_ > 3) {
    fprintf(stderr, "[b200nufft] Invalid type (%d): should be 1, 2, or 3.\n", type_);
    return B2N_ERR_TYPE_NOTVALID;
  }
  if (ntransf_ < 1) {
    fprintf(stderr, "[b200nufft] Invalid ntransf (%d): should be at least 1.\n", ntransf_);
    return B2N_ERR_NTRANS_NOTVALID;
  }
  type = type_;
  dim = dim_;
  iflag = iflag_ >= 0 ? 1 : -1;
  ntransf = ntransf_;
  eps = eps_;
  if (o) opts = *o; else b2n_default_opts(&opts);
  stream = (cudaStream_t)opts.gpu_stream;
  if (opts.gpu_method < 0 || opts.gpu_method > 4) return B2N_ERR_METHOD_NOTVALID;
  if (type == 3) opts.gpu_spreadinterponly = 1;  // outer plan only spreads (impl.h:115-117)
  batch = opts.gpu_maxbatchsize > 0 ? opts.gpu_maxbatchsize : std::min(ntransf, 8);
  batch = std::min(batch, ntransf);
  if (opts.upsampfac == 0.0) {  // auto (impl.h:151-155)
    opts.upsampfac = 2.0;
    if (eps >= 1e-9 && type == 3) opts.upsampfac = 1.25;
  }
  sigma = opts.upsampfac;
  int ier = setup_spreader(eps, sigma, opts.gpu_kerevalmeth, is_double, &ns, &beta);
  if (ier > 1) return ier;
  warn = ier;
  fill_table<T>(tab, ns, beta, opts.gpu_kerevalmeth == 0, &ncoef);

  // method: tile kernels wherever a tile fits, GM otherwise / on request / in 1-D
  method = (opts.gpu_method == 1 || dim == 1) ? 1 : 2;
  if (opts.gpu_binsizex > 0 || opts.gpu_binsizey > 0 || opts.gpu_binsizez > 0) {
    bin[0] = std::max(opts.gpu_binsizex, 1);
    bin[1] = dim > 1 ? std::max(opts.gpu_binsizey, 1) : 1;
    bin[2] = dim > 2 ? std::max(opts.gpu_binsizez, 1) : 1;
    if (method == 2) {
      if (sizeof(T) == 4 && (bin[0] & 1)) return B2N_ERR_BINSIZE_NOTVALID;  // float tiles need even x bins
      size_t b = tile_bytes_for<T>(dim, ns, bin);
      if (b == 0 || b > SMEM_1CTA) {
        if (opts.gpu_method == 0) method = 1;
        else return B2N_ERR_INSUFFICIENT_SHMEM;
      }
    }
  } else {
    if (!choose_bins<T>(dim, ns, bin)) {
      method = 1;
      bin[0] = dim == 1 ? 1024 : 16; bin[1] = dim > 1 ? 16 : 1; bin[2] = dim > 2 ? 4 : 1;
    }
  }
  maxsub = opts.gpu_maxsubprobsize > 0 ? opts.gpu_maxsubprobsize : 2048;
  if (!pts.ovf_host) {  // home of the two-pass sort's overflow flag (sort.cu: learned_skip); optional
    if (cudaHostAlloc(reinterpret_cast<void **>(&pts.ovf_host), sizeof(int), cudaHostAllocDefault) == cudaSuccess) {
      *pts.ovf_host = 0;
    } else {
      pts.ovf_host = nullptr;
      cudaGetLastError();
    }
  }
  base_method = method;
  base_maxsub = maxsub;
  for (int d = 0; d < 3; d++) base_bin[d] = bin[d];
  // sliding-window register kernels: 3-D float, ns <= 8, method auto or 3 ("no shared atomics",
  // the niche of the reference's output-driven method), default bins
  swr_ok = sizeof(T) == 4 && (dim == 3 || dim == 2) && ns <= 8 && method == 2 &&
           (opts.gpu_method == 0 || opts.gpu_method == 3) && opts.gpu_binsizex <= 0 &&
           opts.gpu_binsizey <= 0 && opts.gpu_binsizez <= 0;

  if (type != 3) {
    nmodes = 1;
    for (int d = 0; d < 3; d++) ms[d] = d < dim ? n_modes[d] : 1;
    for (int d = 0; d < dim; d++) nmodes *= ms[d];
    for (int d = 0; d < dim; d++)
      nf[d] = opts.gpu_spreadinterponly ? ms[d] : set_nf_type12(ms[d], sigma, ns);
    if (int e = alloc_grid()) return e;
    if (!opts.gpu_spreadinterponly && !opts.debug && !getenv("B2N_NO_OVERLAP")) {  // side stream of overlap_begin
      int prio_lo = 0, prio_hi = 0;  // highest priority: its blocks go first whenever the sort's leave room
      cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
      if (cudaStreamCreateWithPriority(&side, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_fork, cudaEventDisableTiming) != cudaSuccess ||
          cudaEventCreateWithFlags(&ev_join, cudaEventDisableTiming) != cudaSuccess) {
        cudaGetLastError();
        if (side) cudaStreamDestroy(side);
        side = nullptr;
      }
    }
  }
  if (opts.debug)
    fprintf(stderr, "[b200nufft] plan: type %d dim %d %s eps=%.3g sigma=%.3g ns=%d beta=%.4g ncoef=%d method=%s "
           "bins=(%d,%d,%d) nf=(%ld,%ld,%ld) batch=%d\n",
           type, dim, is_double ? "f64" : "f32", eps, sigma, ns, beta, ncoef, method == 2 ? "tile" : "GM",
           bin[0], bin[1], bin[2], (long)nf[0], (long)nf[1], (long)nf[2], batch);
  return warn;
}

// fine grid, kernel Fourier series and cuFFT plan for the current nf (types 1/2 at plan time,
// type 3 outer/inner at setpts time)
template <typename T> int Plan<T>::alloc_grid() {
  nftot = 1;
  for (int d = 0; d < 3; d++) {
    if (d >= dim) nf[d] = 1;
    nftot *= nf[d];
    nbin[d] = d < dim ? cdiv(nf[d], bin[d]) : 1;
  }
  nbins = (int64_t)nbin[0] * nbin[1] * nbin[2];
  for (int d = 0; d < dim; d++)
    if (nf[d] > 0x7fffffffLL) return B2N_ERR_NDATA_NOTVALID;
  if (nbins > 0x7fffffffLL) return B2N_ERR_NDATA_NOTVALID;
  if (opts.gpu_spreadinterponly) return 0;  // grid is the caller's array; no FFT, no kernel FT
  const int64_t need = (int64_t)batch * nftot;
  if (need > cap_fw) {
    dev_free(fw, stream);
    fw = nullptr;
    cap_fw = 0;
    if (int e = dev_alloc_t(&fw, (size_t)need, stream)) return e;
    cap_fw = need;
  }
  for (int d = 0; d < dim; d++) {
    dev_free(fwker[d], stream);
    fwker[d] = nullptr;
    if (int e = dev_alloc_t(&fwker[d], (size_t)(nf[d] / 2 + 1), stream)) return e;
  }
  if (has_fft) { cufftDestroy(fft); has_fft = false; }
  if (pruned) {
    cufftDestroy(fft_z);
    for (int k = 0; k < 2; k++) if (slab_n[k]) cufftDestroy(fft_xy[k]);
    pruned = false;
  }
  const cufftType ft = sizeof(T) == 4 ? CUFFT_C2C : CUFFT_Z2Z;
  static const char *no_prune = getenv("B2N_NO_PRUNED_FFT");
  if (dim == 3 && type != 3 && !no_prune && 4 * ms[2] <= 3 * nf[2]) {
    // planes that carry modes (deconvolve_wrapper.cu:91-111): [0, (N3-1)/2] and [nf3 - N3/2, nf3)
    slab_lo[0] = 0;               slab_n[0] = (ms[2] - 1) / 2 + 1;
    slab_lo[1] = nf[2] - ms[2] / 2; slab_n[1] = ms[2] / 2;
    const long long plane = nf[0] * nf[1];
    long long nz[1] = {(long long)nf[2]}, nxy[2] = {(long long)nf[1], (long long)nf[0]};
    size_t work = 0;
    bool ok = cufftCreate(&fft_z) == CUFFT_SUCCESS &&
              cufftMakePlanMany64(fft_z, 1, nz, nz, plane, 1, nz, plane, 1, ft, plane, &work) == CUFFT_SUCCESS;
    for (int k = 0; k < 2 && ok; k++) {
      if (!slab_n[k]) continue;
      ok = cufftCreate(&fft_xy[k]) == CUFFT_SUCCESS &&
           cufftMakePlanMany64(fft_xy[k], 2, nxy, nxy, 1, plane, nxy, 1, plane, ft, slab_n[k], &work) == CUFFT_SUCCESS;
    }
    if (ok) {
      pruned = true;
      cufftSetStream(fft_z, stream);
      for (int k = 0; k < 2; k++) if (slab_n[k]) cufftSetStream(fft_xy[k], stream);
      return compute_fseries<T>(*this);
    }
    fprintf(stderr, "[b200nufft] pruned cufft plans failed; using the plain 3-D plan\n");
    if (fft_z) cufftDestroy(fft_z);
    for (int k = 0; k < 2; k++) { if (fft_xy[k]) cufftDestroy(fft_xy[k]); fft_xy[k] = 0; }
    fft_z = 0;
  }
  long long n[3];
  for (int d = 0; d < dim; d++) n[d] = nf[dim - 1 - d];  // slowest first
  if (cufftCreate(&fft) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
  size_t work = 0;
  cufftResult r = cufftMakePlanMany64(fft, dim, n, nullptr, 1, nftot, nullptr, 1, nftot,
                                      sizeof(T) == 4 ? CUFFT_C2C : CUFFT_Z2Z, batch, &work);
  if (r != CUFFT_SUCCESS) {
    fprintf(stderr, "[b200nufft] cufft plan failed (%d)\n", (int)r);
    cufftDestroy(fft);
    return B2N_ERR_CUDA_FAILURE;
  }
  has_fft = true;
  cufftSetStream(fft, stream);
  return compute_fseries<T>(*this);
}

template <typename T>
int Plan<T>::setpts(int64_t M, const void *x, const void *y, const void *z, int64_t N, const void *s,
                    const void *t, const void *u) {
  if (M < 0 || M > 0x7fffffffLL) return B2N_ERR_NDATA_NOTVALID;
  if (type != 3) return setpts12(M, (const T *)x, (const T *)y, (const T *)z);
  return setpts3(M, (const T *)x, (const T *)y, (const T *)z, N, (const T *)s, (const T *)t, (const T *)u);
}

// Geometry that depends on the point count: the SWR kernels pay one pass over a 16 x 12 x (ns+1)
// register window per subproblem, which only amortises on reasonably dense point sets.
template <typename T> void Plan<T>::set_geometry(int64_t M) {
  static const char *force = getenv("B2N_FORCE_METHOD");
  static const double dens3 = getenv("B2N_SWR_DENSITY3") ? atof(getenv("B2N_SWR_DENSITY3")) : 0.05;
  static const double dens2 = getenv("B2N_SWR_DENSITY2") ? atof(getenv("B2N_SWR_DENSITY2")) : 0.10;
  // dense enough to fill the bins (3-D bins hold 2 x 6 x 64 anchor cells, 2-D bins (17 - ns)^2).
  // Thresholds measured on B200: 3-D interp at 0.064 points per cell 1.24 ms (SWR) vs 2.82 ms (tile);
  // 2-D at 0.12 points per cell 0.53 ms (RT2) vs 0.81 ms (tile)
  bool swr = swr_ok && nf[0] % 2 == 0 && nf[0] >= 32 && nf[1] >= 32 &&
             (dim == 3 ? nf[2] >= 32 && (double)M >= dens3 * (double)nftot : (double)M >= dens2 * (double)nftot);
  if (force && swr_ok) swr = force[0] == '3';
  // 2-D type 1 with stacked transforms (jax-finufft's vmap stacking, BASELINE config 4): the
  // narrow-window stacked spreader and its bins
  stacked2 = swr && sizeof(T) == 4 && dim == 2 && type == 1 && batch >= 2 && ns <= 7 && !opts.gpu_spreadinterponly;
  if (swr) {
    method = 3;
    swr_bins(dim, ns, bin, stacked2);
    maxsub = opts.gpu_maxsubprobsize > 0 ? opts.gpu_maxsubprobsize : 2048;
  } else {
    method = base_method;
    maxsub = base_maxsub;
    for (int d = 0; d < 3; d++) bin[d] = base_bin[d];
  }
  for (int d = 0; d < 3; d++) nbin[d] = d < dim ? cdiv(nf[d], bin[d]) : 1;
  nbins = (int64_t)nbin[0] * nbin[1] * nbin[2];
}

template <typename T> int Plan<T>::setpts12(int64_t M, const T *x, const T *y, const T *z) {
  StageTimer tm(opts.debug != 0, stream, &timings[0]);
  set_geometry(M);
  if (opts.debug)
    fprintf(stderr, "[b200nufft] setpts: M=%ld method=%d bins=(%d,%d,%d) maxsub=%d\n", (long)M, method, bin[0],
            bin[1], bin[2], maxsub);
  return binsort_points<T>(*this, M, x, y, z);
}

// Type-3 setpts: V/include/cufinufft/impl.h:461-823
template <typename T>
int Plan<T>::setpts3(int64_t M, const T *x, const T *y, const T *z, int64_t N, const T *s, const T *t,
                     const T *u) {
  if (N < 0) return B2N_ERR_NUM_NU_PTS_INVALID;
  if (N > 0x7fffffffLL) return B2N_ERR_NUM_NU_PTS_INVALID;
  const T *X[3] = {x, y, z}, *S[3] = {s, t, u};
  for (int d = 0; d < dim; d++)
    if (S[d] == nullptr) return B2N_ERR_INVALID_ARGUMENT;
  N3 = N;
  double lohi[12];
  {
    StageTimer tm(opts.debug != 0, stream, &timings[5]);
    if (int e = t3_minmax<T>(stream, dim, M, X, N, S, lohi)) return e;
  }
  for (int d = 0; d < dim; d++) {
    widcen(lohi[2 * d], lohi[2 * d + 1], is_double, &t3X[d], &t3C[d]);
    widcen(lohi[6 + 2 * d], lohi[6 + 2 * d + 1], is_double, &t3S[d], &t3D[d]);
    set_nhg_type3(t3S[d], t3X[d], sigma, ns, is_double, &nf[d], &t3h[d], &t3gam[d]);
  }
  if (opts.debug)
    for (int d = 0; d < dim; d++)
      fprintf(stderr, "[b200nufft] t3 dim %d: X=%.3g C=%.3g S=%.3g D=%.3g gam=%g nf=%ld h=%.3g\n", d, t3X[d],
             t3C[d], t3S[d], t3D[d], t3gam[d], (long)nf[d], t3h[d]);
  // outer grid: spread-only plan geometry + its own fine grid (batch * nf)
  nftot = 1;
  for (int d = 0; d < 3; d++) {
    if (d >= dim) nf[d] = 1;
    nftot *= nf[d];
    nbin[d] = d < dim ? cdiv(nf[d], bin[d]) : 1;
  }
  nbins = (int64_t)nbin[0] * nbin[1] * nbin[2];
  const int64_t need = (int64_t)batch * nftot;
  if (need > cap_fw) {
    dev_free(fw, stream);
    fw = nullptr;
    cap_fw = 0;
    if (int e = dev_alloc_t(&fw, (size_t)need, stream)) return e;
    cap_fw = need;
  }
  if (M > cap_xp) {
    for (int d = 0; d < dim; d++) {
      dev_free(xp[d], stream);
      xp[d] = nullptr;
      if (int e = dev_alloc_t(&xp[d], (size_t)M, stream)) return e;
    }
    dev_free(prephase, stream);
    prephase = nullptr;
    if (int e = dev_alloc_t(&prephase, (size_t)M, stream)) return e;
    cap_xp = M;
  }
  if (N > cap_sp3) {
    for (int d = 0; d < dim; d++) {
      dev_free(sp[d], stream);
      sp[d] = nullptr;
      if (int e = dev_alloc_t(&sp[d], (size_t)N, stream)) return e;
    }
    dev_free(deconv, stream);
    deconv = nullptr;
    if (int e = dev_alloc_t(&deconv, (size_t)N, stream)) return e;
    cap_sp3 = N;
  }
  pts.M = M;
  {
    StageTimer tm(opts.debug != 0, stream, &timings[5]);
    if (int e = t3_prepare<T>(*this, X, S)) return e;
  }
  // bin-sort the rescaled sources for the outer spread
  if (int e = setpts12(M, xp[0], xp[1], xp[2])) return e;
  // inner type-2 plan: modes = outer fine grid, modeord 0, same iflag/eps/sigma (impl.h:795-812)
  b2n_opts io = opts;
  io.gpu_spreadinterponly = 0;
  io.gpu_method = 0;
  io.modeord = 0;
  io.gpu_maxbatchsize = batch;
  io.gpu_stream = stream;
  bool rebuild = inner == nullptr;
  if (inner)
    for (int d = 0; d < dim; d++)
      if (inner->ms[d] != nf[d]) rebuild = true;
  if (rebuild) {
    delete inner;
    inner = new (std::nothrow) Plan<T>();
    if (!inner) return B2N_ERR_ALLOC;
    int e = inner->init(2, dim, nf, iflag, batch, eps, &io);
    if (e > 1) {
      delete inner;
      inner = nullptr;
      return e;
    }
  }
  // the rescaled targets |s'| <= pi / sigma fill the central 1 / sigma of the inner grid's axes
  // (impl.h:700-712: s' = h gamma (s - D), gamma = nf / (2 sigma S))
  inner->sort_fill = 1.0 / inner->sigma;
  return inner->setpts12(N, sp[0], sp[1], sp[2]);
}

template <typename T> int Plan<T>::execute(void *c, void *fk) {
  if (type == 1) return exec1((cpx<T> *)c, (cpx<T> *)fk);
  if (type == 2) return exec2((cpx<T> *)c, (cpx<T> *)fk, nullptr);
  return exec3((cpx<T> *)c, (cpx<T> *)fk);
}

// Chunked execution (b2n_run_host).  Spreading is additive over point subsets and interpolation
// is independent per point, so a transform over M points equals the same transform over its
// chunks with the uniform-grid stages done once.
template <typename T> int Plan<T>::exec_phase(int phase, void *cv, void *fkv) {
  if (type == 3 || ntransf > batch || opts.gpu_spreadinterponly) return B2N_ERR_METHOD_NOTVALID;
  const bool dbg = opts.debug != 0;
  cpx<T> *c = (cpx<T> *)cv, *fk = (cpx<T> *)fkv;
  if (type == 1) {
    if (phase & PH_BEGIN) {
      StageTimer tm(dbg, stream, &timings[6]);
      B2N_CUDA_OK(cudaMemsetAsync(fw, 0, sizeof(cpx<T>) * (size_t)ntransf * nftot, stream));
    }
    if (phase & PH_BODY) {
      StageTimer tm(dbg, stream, &timings[1]);
      if (int e = spread(c, nullptr, fw, ntransf)) return e;
    }
    if (phase & PH_END) {
      {
        StageTimer tm(dbg, stream, &timings[2]);
        if (int e = run_fft(*this, ntransf)) return e;
      }
      StageTimer tm(dbg, stream, &timings[3]);
      if (int e = deconvolve<T>(*this, fw, fk, ntransf)) return e;
    }
    return 0;
  }
  if (phase & PH_BEGIN) {
    {
      StageTimer tm(dbg, stream, &timings[3]);
      if (int e = amplify<T>(*this, fw, fk, ntransf)) return e;
    }
    StageTimer tm(dbg, stream, &timings[2]);
    if (int e = run_fft(*this, ntransf)) return e;
  }
  if (phase & PH_BODY) {
    StageTimer tm(dbg, stream, &timings[4]);
    if (int e = interp(c, nullptr, fw, ntransf)) return e;
  }
  return 0;
}

template <typename T> bool Plan<T>::overlap_begin(void *c, void *fk, cudaEvent_t wait_first, int *err) {
  *err = 0;
  // worth a fork only when the grid stages are not tiny (two event operations + a stream switch)
  if (!side || type == 3 || ntransf > batch || opts.gpu_spreadinterponly || opts.debug ||
      (size_t)nftot * sizeof(cpx<T>) < (size_t(32) << 20))
    return false;
  cudaStream_t main_st = stream;
  if (cudaEventRecord(ev_fork, main_st) != cudaSuccess || cudaStreamWaitEvent(side, ev_fork, 0) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  if (wait_first && type == 2) cudaStreamWaitEvent(side, wait_first, 0);  // the modes, when they come from the host
  set_stream(side);
  *err = exec_phase(PH_BEGIN, c, fk);
  set_stream(main_st);
  if (cudaEventRecord(ev_join, side) != cudaSuccess) *err = *err ? *err : B2N_ERR_CUDA_FAILURE;
  return true;
}
template <typename T> void Plan<T>::overlap_join() { cudaStreamWaitEvent(stream, ev_join, 0); }

// Chunk geometry for a host-resident point set: enough chunks that the copy of chunk k+1 hides
// the bin-sort + spread/interp of chunk k, few enough that a chunk still fills the bins (the
// sliding-window kernels pay one pass over every bin's window per point set).
template <typename T> bool Plan<T>::can_chunk(int64_t M, int *nchunk, int64_t *chunk) const {
  static const char *off = getenv("B2N_NO_CHUNK");
  if (off || type == 3 || ntransf != 1 || opts.gpu_spreadinterponly || M < (int64_t(1) << 23)) return false;
  int64_t n = std::min<int64_t>(8, M / std::max<int64_t>(1, (int64_t)(0.06 * (double)nftot)));
  if (n < 2) return false;
  int64_t ch = ((M + n - 1) / n + 31) & ~int64_t(31);  // keeps every chunk's arrays 128-byte aligned
  *chunk = ch;
  *nchunk = (int)((M + ch - 1) / ch);
  return *nchunk >= 2;
}

// The float tile kernels flush / gather cell PAIRS with 16-byte accesses (tile_kernels.cuh: red16,
// float4 loads): every row and every transform's grid must start on a 16-byte boundary.  Our own
// fine grids always do (nf even); with gpu_spreadinterponly the grid is the caller's array
// (nf = n_modes), which may have an odd row length, an odd size per transform, or an unaligned
// base: those run on the GM kernels (scalar 8-byte accesses) instead.
template <typename T> static bool tile_grid_ok(const Plan<T> &p, const void *grid, int ntr) {
  if (sizeof(T) != 4) return true;
  return ((uintptr_t)grid & 15) == 0 && (p.nf[0] & 1) == 0 && (ntr <= 1 || (p.nftot & 1) == 0);
}
template <typename T>
int Plan<T>::spread(const cpx<T> *c, const cpx<T> *prescale, cpx<T> *grid, int ntr) {
  if constexpr (sizeof(T) == 4) {  // ns = 8: two cells (16 bytes) per lane
    if (method == 3 && (ns < 8 || dim == 2 || tile_grid_ok(*this, grid, ntr))) return spread_swr(*this, c, prescale, grid, ntr);
  }
  return method == 2 && tile_grid_ok(*this, grid, ntr) ? spread_tile<T>(*this, c, prescale, grid, ntr)
                                                       : spread_gm<T>(*this, c, prescale, grid, ntr);
}
template <typename T>
int Plan<T>::interp(cpx<T> *c, const cpx<T> *postscale, const cpx<T> *grid, int ntr) {
  if constexpr (sizeof(T) == 4) {
    if (method == 3 && (ns < 8 || dim == 2 || tile_grid_ok(*this, grid, ntr))) return interp_swr(*this, c, postscale, grid, ntr);
  }
  return method == 2 && tile_grid_ok(*this, grid, ntr) ? interp_tile<T>(*this, c, postscale, grid, ntr)
                                                       : interp_gm<T>(*this, c, postscale, grid, ntr);
}

template <typename T> static cufftResult fft_exec(cufftHandle h, cpx<T> *d, int dir) {
  if (sizeof(T) == 4) return cufftExecC2C(h, (cufftComplex *)d, (cufftComplex *)d, dir);
  return cufftExecZ2Z(h, (cufftDoubleComplex *)d, (cufftDoubleComplex *)d, dir);
}

// blk = transforms of the current batch that hold data
template <typename T> static int run_fft(Plan<T> &p, int blk = -1, cpx<T> *grid = nullptr) {
  const int dir = p.iflag >= 0 ? CUFFT_INVERSE : CUFFT_FORWARD;  // cufft_ex(.., iflag): types.h:108-115
  if (p.pruned) {
    // type 1: the grid is full, only the mode planes of the result are read  -> z first, then
    //         (y, x) on the kept planes;  type 2: only the mode planes are non-zero -> (y, x) on
    //         them first, then z everywhere.  (A DFT is separable: any order gives the same sum.)
    const int nb = blk < 0 ? p.batch : blk;
    const int64_t plane = p.nf[0] * p.nf[1];
    for (int t = 0; t < nb; t++) {
      cpx<T> *g = p.fw + (int64_t)t * p.nftot;
      if (p.type == 1 && fft_exec<T>(p.fft_z, g, dir) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
      for (int k = 0; k < 2; k++)
        if (p.slab_n[k] && fft_exec<T>(p.fft_xy[k], g + p.slab_lo[k] * plane, dir) != CUFFT_SUCCESS)
          return B2N_ERR_CUDA_FAILURE;
      if (p.type != 1 && fft_exec<T>(p.fft_z, g, dir) != CUFFT_SUCCESS) return B2N_ERR_CUDA_FAILURE;
    }
    return 0;
  }
  return fft_exec<T>(p.fft, grid ? grid : p.fw, dir) == CUFFT_SUCCESS ? 0 : B2N_ERR_CUDA_FAILURE;
}

// Stacked 2-D type 1 (jax-finufft's vmap stacking, BASELINE config 4) with more transforms than one
// batch: the spreader takes a whole GROUP of batches in one launch (rt2_kernels.cuh: k_rt2s_spread
// runs its passes of 8 transforms back to back, kernel vectors evaluated once), into a stack of
// fine grids of its own; FFT and deconvolution then go through the stack batch by batch with the
// plan's batched cuFFT handle.  The group is bounded by B2N_STACK_BYTES (default 8 GB) for the
// stack plus the point-major copy of the strengths.
template <typename T> int Plan<T>::exec1_stacked(cpx<T> *c, cpx<T> *fk, int group) {
  const bool dbg = opts.debug != 0;
  const int64_t M = pts.M;
  const int64_t need = (int64_t)(group + batch) * nftot;  // + one batch: the cuFFT handle always runs `batch` grids
  if (need > cap_fw_stack) {
    dev_free(fw_stack, stream);
    fw_stack = nullptr;
    cap_fw_stack = 0;
    if (int e = dev_alloc_t(&fw_stack, (size_t)need, stream)) return e;
    cap_fw_stack = need;
  }
  for (int g0 = 0; g0 < ntransf; g0 += group) {
    const int ng = std::min(group, ntransf - g0);
    {
      StageTimer tm(dbg, stream, &timings[6]);
      const size_t nclear = (size_t)((ng + batch - 1) / batch * batch);  // whole batches: the FFT runs on the padding too
      B2N_CUDA_OK(cudaMemsetAsync(fw_stack, 0, sizeof(cpx<T>) * nclear * (size_t)nftot, stream));
    }
    {
      StageTimer tm(dbg, stream, &timings[1]);
      if (int e = spread(c + (int64_t)g0 * M, nullptr, fw_stack, ng)) return e;
    }
    for (int b = 0; b < ng; b += batch) {
      const int blk = std::min(batch, ng - b);
      cpx<T> *grid = fw_stack + (int64_t)b * nftot;
      {
        StageTimer tm(dbg, stream, &timings[2]);
        if (int e = run_fft(*this, blk, grid)) return e;
      }
      StageTimer tm(dbg, stream, &timings[3]);
      if (int e = deconvolve<T>(*this, grid, fk + (int64_t)(g0 + b) * nmodes, blk)) return e;
    }
  }
  return 0;
}

// Type 1: V/src/cuda/3d/cufinufft3d.cu:18-73 (and 1d/2d twins)
template <typename T> int Plan<T>::exec1(cpx<T> *c, cpx<T> *fk) {
  const bool dbg = opts.debug != 0;
  const int64_t M = pts.M;
  if (stacked2 && method == 3 && ntransf > batch && batch == 8 && !pruned) {
    static const double budget = getenv("B2N_STACK_BYTES") ? atof(getenv("B2N_STACK_BYTES")) : 8e9;
    const int64_t g = (int64_t)(budget / (sizeof(cpx<T>) * (double)(nftot + M))) / batch * batch;
    if (g >= 2 * batch) return exec1_stacked(c, fk, (int)std::min<int64_t>(g, (ntransf + batch - 1) / batch * batch));
  }
  for (int i = 0; i * batch < ntransf; i++) {
    const int blk = std::min(ntransf - i * batch, batch);
    cpx<T> *cs = c + (int64_t)i * batch * M;
    cpx<T> *fks = fk + (int64_t)i * batch * nmodes;
    cpx<T> *grid = opts.gpu_spreadinterponly ? fks : fw;
    {
      StageTimer tm(dbg, stream, &timings[6]);
      B2N_CUDA_OK(cudaMemsetAsync(grid, 0, sizeof(cpx<T>) * (size_t)blk * nftot, stream));
    }
    {
      StageTimer tm(dbg, stream, &timings[1]);
      if (int e = spread(cs, nullptr, grid, blk)) return e;
    }
    if (opts.gpu_spreadinterponly) continue;
    {
      StageTimer tm(dbg, stream, &timings[2]);
      if (int e = run_fft(*this, blk)) return e;
    }
    {
      StageTimer tm(dbg, stream, &timings[3]);
      if (int e = deconvolve<T>(*this, fw, fks, blk)) return e;
    }
  }
  return 0;
}

// Type 2: V/src/cuda/3d/cufinufft3d.cu:75-125
template <typename T> int Plan<T>::exec2(cpx<T> *c, cpx<T> *fk, const cpx<T> *postscale) {
  const bool dbg = opts.debug != 0;
  const int64_t M = pts.M;
  for (int i = 0; i * batch < ntransf; i++) {
    const int blk = std::min(ntransf - i * batch, batch);
    cpx<T> *cs = c + (int64_t)i * batch * M;
    cpx<T> *fks = fk + (int64_t)i * batch * nmodes;
    const cpx<T> *grid = fks;
    if (!opts.gpu_spreadinterponly) {
      {
        StageTimer tm(dbg, stream, &timings[3]);
        if (int e = amplify<T>(*this, fw, fks, blk)) return e;
      }
      {
        StageTimer tm(dbg, stream, &timings[2]);
        if (int e = run_fft(*this, blk)) return e;
      }
      grid = fw;
    }
    {
      StageTimer tm(dbg, stream, &timings[4]);
      if (int e = interp(cs, postscale, grid, blk)) return e;
    }
  }
  return 0;
}

// Type 3: V/src/cuda/3d/cufinufft3d.cu:127-183.  The reference makes two extra passes per
// transform (prephase*c into CpBatch; deconv*f in place); here the prephase multiply is fused
// into the spreader's strength load and the deconv multiply into the interpolator's store.
template <typename T> int Plan<T>::exec3(cpx<T> *c, cpx<T> *fk) {
  const bool dbg = opts.debug != 0;
  const int64_t M = pts.M;
  bool anyD = false;
  for (int d = 0; d < dim; d++) anyD = anyD || t3D[d] != 0;
  for (int i = 0; i * batch < ntransf; i++) {
    const int blk = std::min(ntransf - i * batch, batch);
    cpx<T> *cs = c + (int64_t)i * batch * M;
    cpx<T> *fks = fk + (int64_t)i * batch * N3;
    {
      StageTimer tm(dbg, stream, &timings[6]);
      B2N_CUDA_OK(cudaMemsetAsync(fw, 0, sizeof(cpx<T>) * (size_t)blk * nftot, stream));
    }
    {
      StageTimer tm(dbg, stream, &timings[1]);
      if (int e = spread(cs, anyD ? prephase : nullptr, fw, blk)) return e;
    }
    inner->ntransf = blk;
    if (int e = inner->exec2(fks, fw, deconv)) return e;
    if (dbg)
      for (int k = 2; k <= 4; k++) { timings[k] += inner->timings[k]; inner->timings[k] = 0; }
  }
  return 0;
}

template <typename T> void Plan<T>::info(b2n_plan_info *o) {
  std::memset(o, 0, sizeof(*o));
  o->type = type; o->dim = dim; o->is_double = is_double; o->ns = ns; o->method = method;
  o->ntransf = ntransf; o->batchsize = batch; o->ncoef = ncoef; o->beta = beta; o->upsampfac = sigma;
  for (int d = 0; d < 3; d++) {
    o->nf[d] = nf[d]; o->ms[d] = ms[d]; o->binsize[d] = bin[d]; o->nbins[d] = nbin[d];
    o->t3_X[d] = t3X[d]; o->t3_C[d] = t3C[d]; o->t3_S[d] = t3S[d]; o->t3_D[d] = t3D[d];
    o->t3_h[d] = t3h[d]; o->t3_gam[d] = t3gam[d];
    o->t3_nf_inner[d] = inner ? inner->nf[d] : 0;
  }
  o->M = pts.M;
  o->N = N3;
}

template <typename T>
int Plan<T>::sort_get(const int32_t **idx, const int32_t **bin_start, int64_t *nb) {
  if (int e = materialise_idx<T>(*this)) return e;
  *idx = pts.idx;
  *bin_start = pts.bin_start;
  *nb = nbins;
  return 0;
}

template struct Plan<float>;
template struct Plan<double>;

// ------------------------------------------------------------------------------------ plan cache
// run_nufft in the reference builds and destroys a plan (cuFFT plan, ~12 allocations, kernel FT)
// on EVERY custom call (lib/kernels.cc.cu:49-51,84).  Here finished plans are parked, keyed by
// everything that shapes them, and re-used by the next identical call; an event orders re-use
// across streams.  Thread-safe: a plan is owned by exactly one caller while checked out.
struct CacheKey {
  int type, dim, is_double, iflag, ntransf, device;
  int64_t nk[3];
  double eps, upsampfac;
  int modeord, method, sort, kerevalmeth, maxbatch, debug;
  bool operator==(const CacheKey &o) const { return std::memcmp(this, &o, sizeof(CacheKey)) == 0; }
};
struct CacheEntry {
  CacheKey key;
  PlanBase *plan;
  cudaEvent_t done;            // last use of the plan (null for graph-owned entries)
  unsigned long long owner;    // 0, or the id of the stream capture this plan now belongs to
};
static std::mutex g_mu;
static std::vector<CacheEntry> g_cache;   // free plans, oldest first
static std::vector<CacheEntry> g_pinned;  // plans recorded into a CUDA graph (see cache_put)
constexpr size_t CACHE_MAX = 8;
// ... and by bytes: parked plans keep their fine grids, sort workspaces and type-3 arrays (a 3-D
// grid with 8 stacked transforms is several GB per key), invisible to the host framework's own
// allocator.  After a plan is parked the oldest ones are dropped until the pool's used bytes fit
// the limit (default: a quarter of the device; B2N_CACHE_BYTES / b2n_set_cache_limit).
static std::atomic<long long> g_cache_limit{-1};
static size_t cache_limit_bytes() {
  long long v = g_cache_limit.load(std::memory_order_relaxed);
  if (v < 0) {
    const char *e = getenv("B2N_CACHE_BYTES");
    if (e && e[0]) v = atoll(e);
    else {
      size_t fr = 0, tot = 0;
      v = cudaMemGetInfo(&fr, &tot) == cudaSuccess ? (long long)(tot / 4) : (16LL << 30);
    }
    g_cache_limit.store(v, std::memory_order_relaxed);
  }
  return (size_t)v;
}

// CUDA-graph capture.  A call made while its stream is capturing records the
// plan's kernels, and with them the plan's buffers, into the graph: from then on the plan
// belongs to that graph.  It is parked under the capture's id, handed out again only to calls of
// the same capture (in-stream order protects it there) and never to eager callers; it lives
// until b2n_cache_clear().  Eager calls after the capture build / take other plans.
static unsigned long long capture_id(cudaStream_t st) {
  cudaStreamCaptureStatus status = cudaStreamCaptureStatusNone;
  unsigned long long id = 0;
  if (cudaStreamGetCaptureInfo(st, &status, &id) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return status == cudaStreamCaptureStatusActive ? (id ? id : 1) : 0;
}
// Host-side CUDA calls that are "potentially unsafe" while some stream of the process captures in
// global mode (event synchronisation, cudaMalloc inside cufftPlanMany): made in relaxed mode.
struct RelaxedCapture {
  cudaStreamCaptureMode mode = cudaStreamCaptureModeRelaxed;
  bool on;
  explicit RelaxedCapture(bool enable) : on(enable) { if (on) cudaThreadExchangeStreamCaptureMode(&mode); }
  ~RelaxedCapture() { if (on) cudaThreadExchangeStreamCaptureMode(&mode); }
};

static PlanBase *cache_take(const CacheKey &k, cudaStream_t st, unsigned long long cid) {
  CacheEntry e;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (cid)
      for (size_t i = 0; i < g_pinned.size(); i++)
        if (g_pinned[i].owner == cid && g_pinned[i].key == k) {
          PlanBase *p = g_pinned[i].plan;
          g_pinned.erase(g_pinned.begin() + i);
          return p;  // same capture: ordered by the capture's own dependencies
        }
    size_t i = 0;
    for (; i < g_cache.size(); i++)
      if (g_cache[i].key == k) break;
    if (i == g_cache.size()) return nullptr;
    e = g_cache[i];
    g_cache.erase(g_cache.begin() + i);
  }
  if (cid) {  // a dependency on an event outside the capture cannot be recorded: wait on the host
    RelaxedCapture rc(true);
    cudaEventSynchronize(e.done);
  } else {
    cudaStreamWaitEvent(st, e.done, 0);
  }
  cudaEventDestroy(e.done);
  return e.plan;
}
static void cache_put(const CacheKey &k, PlanBase *p, cudaStream_t st, unsigned long long cid) {
  CacheEntry e;
  e.key = k;
  e.plan = p;
  e.done = nullptr;
  e.owner = cid;
  if (cid) {
    std::lock_guard<std::mutex> lk(g_mu);
    g_pinned.push_back(e);
    return;
  }
  cudaEventCreateWithFlags(&e.done, cudaEventDisableTiming);
  cudaEventRecord(e.done, st);
  PlanBase *evict = nullptr;
  cudaEvent_t evict_ev = nullptr;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    if (g_cache.size() >= CACHE_MAX) {
      evict = g_cache.front().plan;
      evict_ev = g_cache.front().done;
      g_cache.erase(g_cache.begin());
    }
    g_cache.push_back(e);
  }
  if (evict) {
    cudaEventSynchronize(evict_ev);
    cudaEventDestroy(evict_ev);
    delete evict;
  }
  // byte bound: drop the oldest parked plans (never the one just parked) while over the limit
  for (;;) {
    size_t reserved = 0, used = 0;
    pool_usage(&reserved, &used);
    if (used <= cache_limit_bytes()) break;
    CacheEntry old;
    {
      std::lock_guard<std::mutex> lk(g_mu);
      if (g_cache.size() <= 1) break;
      old = g_cache.front();
      g_cache.erase(g_cache.begin());
    }
    cudaEventSynchronize(old.done);
    cudaEventDestroy(old.done);
    delete old.plan;
    cudaStreamSynchronize(st);  // the frees above are stream-ordered: make them count before re-checking
  }
}

}  // namespace b2n

// =========================================================================================== C ABI
using namespace b2n;

extern "C" {

const char *b2n_version(void) { return "b200nufft 0.1 (sm_100a)"; }

unsigned long long b2n_launch_count(void) { return g_launch_count.load(std::memory_order_relaxed); }

int b2n_set_setpts_cache(int on) {
  const int prev = setpts_cache_enabled() ? 1 : 0;
  g_setpts_cache.store(on ? 1 : 0, std::memory_order_relaxed);
  return prev;
}

void b2n_default_opts(b2n_opts *o) {  // defaults of V/src/cuda/cufinufft.cu:133-152
  std::memset(o, 0, sizeof(*o));
  o->modeord = 0;
  o->upsampfac = 0.0;
  o->gpu_method = 0;
  o->gpu_sort = 1;
  o->gpu_kerevalmeth = 1;
  o->gpu_maxbatchsize = 0;
  o->debug = 0;
  o->gpu_maxsubprobsize = 0;
  o->gpu_spreadinterponly = 0;
  o->gpu_device_id = 0;
  o->gpu_stream = nullptr;
}

static bool invalid_modes(int type, int dim, const int64_t *m) {  // cufinufft.cu:12-29
  if (type == 3) return false;
  int64_t tot = 1;
  for (int i = 0; i < dim; i++) {
    if (m[i] > 0x7fffffffLL || m[i] <= 0) return true;
    tot *= m[i];
    if (tot > 0x7fffffffLL) return true;
  }
  return false;
}

int b2n_makeplan(int type, int dim, const int64_t *n_modes, int iflag, int ntransf, double eps,
                 int is_double, b2n_plan *plan, const b2n_opts *opts) {
  *plan = nullptr;
  if (dim < 1 || dim > 3) {
    fprintf(stderr, "[b200nufft] Invalid dim (%d), should be 1, 2 or 3.\n", dim);
    return B2N_ERR_DIM_NOTVALID;
  }
  if (type != 3 && (!n_modes || invalid_modes(type, dim, n_modes))) return B2N_ERR_NDATA_NOTVALID;
  PlanBase *p = nullptr;
  int ier;
  try {
    if (is_double) {
      auto *q = new Plan<double>();
      p = q;
      ier = q->init(type, dim, n_modes, iflag, ntransf, eps, opts);
    } else {
      auto *q = new Plan<float>();
      p = q;
      ier = q->init(type, dim, n_modes, iflag, ntransf, eps, opts);
    }
  } catch (...) {
    delete p;
    return B2N_ERR_ALLOC;
  }
  if (ier > 1) {
    delete p;
    return ier;
  }
  *plan = reinterpret_cast<b2n_plan>(p);
  return ier;
}

int b2n_setpts(b2n_plan plan, int64_t M, const void *x, const void *y, const void *z, int64_t N,
               const void *s, const void *t, const void *u) {
  if (!plan) return B2N_ERR_PLAN_NOTVALID;
  try {
    return reinterpret_cast<PlanBase *>(plan)->setpts(M, x, y, z, N, s, t, u);
  } catch (...) {
    return B2N_ERR_ALLOC;
  }
}

int b2n_execute(b2n_plan plan, void *c, void *fk) {
  if (!plan) return B2N_ERR_PLAN_NOTVALID;
  try {
    return reinterpret_cast<PlanBase *>(plan)->execute(c, fk);
  } catch (...) {
    return B2N_ERR_CUDA_FAILURE;
  }
}

int b2n_destroy(b2n_plan plan) {
  if (!plan) return B2N_ERR_PLAN_NOTVALID;  // cufinufft_destroy(NULL) -> 16
  delete reinterpret_cast<PlanBase *>(plan);
  return 0;
}

int b2n_plan_info_get(b2n_plan plan, b2n_plan_info *info) {
  if (!plan) return B2N_ERR_PLAN_NOTVALID;
  reinterpret_cast<PlanBase *>(plan)->info(info);
  return 0;
}

int b2n_plan_sort_get(b2n_plan plan, const int32_t **idx, const int32_t **bin_start, int64_t *nbins) {
  if (!plan) return B2N_ERR_PLAN_NOTVALID;
  return reinterpret_cast<PlanBase *>(plan)->sort_get(idx, bin_start, nbins);
}

int b2n_plan_sort_copy(b2n_plan plan, int32_t *idx_out, int32_t *bin_start_out) {
  if (!plan) return B2N_ERR_PLAN_NOTVALID;
  PlanBase *p = reinterpret_cast<PlanBase *>(plan);
  const int32_t *idx = nullptr, *bs = nullptr;
  int64_t nb = 0;
  if (int e = p->sort_get(&idx, &bs, &nb)) return e;
  b2n_plan_info inf;
  p->info(&inf);
  cudaStream_t st = p->get_stream();
  B2N_CUDA_OK(cudaMemcpyAsync(idx_out, idx, sizeof(int32_t) * (size_t)inf.M, cudaMemcpyDeviceToDevice, st));
  B2N_CUDA_OK(cudaMemcpyAsync(bin_start_out, bs, sizeof(int32_t) * (size_t)(nb + 1), cudaMemcpyDeviceToDevice, st));
  B2N_CUDA_OK(cudaStreamSynchronize(st));
  return 0;
}

int b2n_plan_timings(b2n_plan plan, double *ms7) {
  if (!plan) return B2N_ERR_PLAN_NOTVALID;
  PlanBase *p = reinterpret_cast<PlanBase *>(plan);
  for (int i = 0; i < 7; i++) { ms7[i] = p->timings[i]; p->timings[i] = 0; }
  return 0;
}

void b2n_cache_clear(void) {
  std::vector<CacheEntry> old, pinned;
  {
    std::lock_guard<std::mutex> lk(g_mu);
    old.swap(g_cache);
    pinned.swap(g_pinned);
  }
  for (auto &e : old) {
    cudaEventSynchronize(e.done);
    cudaEventDestroy(e.done);
    delete e.plan;
  }
  if (!pinned.empty()) cudaDeviceSynchronize();  // graphs that recorded these plans must be gone by now
  for (auto &e : pinned) delete e.plan;
  cudaDeviceSynchronize();  // the frees are stream-ordered
  pool_trim(0);             // ... and the memory goes back to the driver
}

long long b2n_set_cache_limit(long long bytes) {
  const long long prev = (long long)cache_limit_bytes();
  g_cache_limit.store(bytes < 0 ? -1 : bytes, std::memory_order_relaxed);
  return prev;
}

void b2n_cache_bytes(unsigned long long *reserved, unsigned long long *used) {
  size_t r = 0, u = 0;
  pool_usage(&r, &u);
  if (reserved) *reserved = r;
  if (used) *used = u;
}

// The body of b2n_run.  src_ready (optional): an event the stream must wait for before the first
// execute reads `src` -- lets b2n_run_host overlap the strengths' H2D copy with the bin-sort.
struct ChunkCtx {  // filled by b2n_run_host when the inputs arrive in chunks
  // plan known: decides the chunking and enqueues the copies (returns 0, or an error code)
  int (*enqueue)(ChunkCtx *, PlanBase *) = nullptr;
  void *user = nullptr;
  int nchunk = 0;
  int64_t chunk = 0;
  cudaEvent_t *pts_ev = nullptr, *src_ev = nullptr;  // per chunk: coordinates / strengths arrived
  cudaEvent_t *out_ev = nullptr;                      // type 2: chunk's outputs are ready on `stream`
  cudaStream_t d2h = nullptr;
  void *host_out = nullptr;
};

static int run_core(int type, int dim, int is_double, cudaStream_t stream, double eps, int iflag, int64_t n_tot,
                    int n_transf, int64_t n_j, const int64_t *n_k, const b2n_opts *opts_in, const void *src,
                    const void *const *pts, const void *const *tgt, void *out, cudaEvent_t src_ready,
                    ChunkCtx *cc = nullptr) {
  b2n_opts o;
  if (opts_in) o = *opts_in; else b2n_default_opts(&o);
  o.gpu_stream = stream;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B2N_ERR_CUDA_FAILURE;  // kernels.cc.cu:40-47
  o.gpu_device_id = dev;
  if (dim < 1 || dim > 3) return B2N_ERR_DIM_NOTVALID;
  if (type < 1 || type > 3) return B2N_ERR_TYPE_NOTVALID;

  CacheKey key;
  std::memset(&key, 0, sizeof(key));
  key.type = type; key.dim = dim; key.is_double = is_double; key.iflag = iflag >= 0 ? 1 : -1;
  key.ntransf = n_transf; key.device = dev;
  for (int d = 0; d < 3; d++) key.nk[d] = (type != 3 && d < dim) ? n_k[d] : 0;
  key.eps = eps; key.upsampfac = o.upsampfac; key.modeord = o.modeord; key.method = o.gpu_method;
  key.sort = o.gpu_sort; key.kerevalmeth = o.gpu_kerevalmeth; key.maxbatch = o.gpu_maxbatchsize;
  key.debug = o.debug;

  const unsigned long long cid = capture_id(stream);
  PlanBase *p = cache_take(key, stream, cid);
  int warn = 0;
  if (p) {
    p->set_stream(stream);
    warn = p->plan_warning;
  } else {
    b2n_plan h = nullptr;
    int64_t nk3[3] = {n_k ? n_k[0] : 1, n_k ? n_k[1] : 1, n_k ? n_k[2] : 1};
    RelaxedCapture rc(cid != 0);  // cufftPlanMany allocates; the plan's own work is stream-ordered
    int ier = b2n_makeplan(type, dim, nk3, iflag, n_transf, eps, is_double, &h, &o);
    if (ier > 1) return ier;  // ret == 1 is a warning (kernels.cc.cu:52)
    warn = ier;
    p = reinterpret_cast<PlanBase *>(h);
    p->plan_warning = warn;
  }
  int64_t n_k_total = 1;
  if (type != 3) for (int d = 0; d < dim; d++) n_k_total *= n_k[d];
  else n_k_total = n_k[0];
  const size_t rs = is_double ? 8 : 4, cs = 2 * rs;
  const int64_t n_src = type == 2 ? n_k_total : n_j;   // per-transform source length
  const int64_t n_out = type == 2 ? n_j : n_k_total;   // per-transform output length
  if (cc) {  // host-resident inputs: let the caller enqueue its copies now that the plan is known
    if (int e = cc->enqueue(cc, p)) {
      cudaStreamSynchronize(stream);
      delete p;
      return e;
    }
  }
  if (cc && cc->nchunk >= 2) {
    int ret = 0;
    for (int k = 0; k < cc->nchunk && ret == 0; k++) {
      const int64_t off = (int64_t)k * cc->chunk, m = std::min<int64_t>(cc->chunk, n_j - off);
      const void *P[3] = {nullptr, nullptr, nullptr};
      for (int d = 0; d < dim; d++) P[d] = (const char *)pts[d] + (size_t)off * rs;
      cudaStreamWaitEvent(stream, cc->pts_ev[k], 0);
      ret = p->setpts(m, P[0], P[1], P[2], 0, nullptr, nullptr, nullptr);
      if (ret) break;
      const int ph = (k == 0 ? PlanBase::PH_BEGIN : 0) | PlanBase::PH_BODY |
                     (k == cc->nchunk - 1 ? PlanBase::PH_END : 0);
      if (type == 1) {
        cudaStreamWaitEvent(stream, cc->src_ev[k], 0);
        ret = p->exec_phase(ph, (char *)src + (size_t)off * cs, out);
      } else {
        if (k == 0) cudaStreamWaitEvent(stream, cc->src_ev[0], 0);  // the modes
        ret = p->exec_phase(ph, (char *)out + (size_t)off * cs, (void *)src);
        if (ret == 0) {  // this chunk's outputs go home while the next chunk is interpolated
          cudaEventRecord(cc->out_ev[k], stream);
          cudaStreamWaitEvent(cc->d2h, cc->out_ev[k], 0);
          cudaMemcpyAsync((char *)cc->host_out + (size_t)off * cs, (char *)out + (size_t)off * cs,
                          (size_t)m * cs, cudaMemcpyDeviceToHost, cc->d2h);
        }
      }
    }
    if (ret != 0) {
      cudaStreamSynchronize(stream);
      if (cc->d2h) cudaStreamSynchronize(cc->d2h);
      delete p;
      return ret;
    }
    n_tot = 0;  // done: skip the whole-array loop below
  }
  for (int64_t index = 0; index < n_tot; index++) {
    const void *P[3] = {nullptr, nullptr, nullptr}, *Tg[3] = {nullptr, nullptr, nullptr};
    for (int d = 0; d < dim; d++) {
      P[d] = (const char *)pts[d] + (size_t)index * n_j * rs;
      if (type == 3) Tg[d] = (const char *)tgt[d] + (size_t)index * n_k_total * rs;
    }
    const char *s_i = (const char *)src + (size_t)index * n_src * n_transf * cs;
    char *o_i = (char *)out + (size_t)index * n_out * n_transf * cs;
    // c = nonuniform side, fk = uniform side (or type-3 targets)
    void *c_i = type == 2 ? (void *)o_i : (void *)s_i, *fk_i = type == 2 ? (void *)s_i : (void *)o_i;
    // the grid stages that do not need the points run beside the bin-sort (PlanBase::overlap_begin)
    int oerr = 0;
    const bool overlapped = p->overlap_begin(c_i, fk_i, index == 0 ? src_ready : nullptr, &oerr);
    int ret = oerr ? oerr : p->setpts(n_j, P[0], P[1], P[2], type == 3 ? n_k_total : 0, Tg[0], Tg[1], Tg[2]);
    if (overlapped) p->overlap_join();
    if (ret != 0) {
      cudaStreamSynchronize(stream);
      delete p;
      return ret;
    }
    if (src_ready && index == 0) cudaStreamWaitEvent(stream, src_ready, 0);
    if (overlapped) ret = p->exec_phase(PlanBase::PH_BODY | PlanBase::PH_END, c_i, fk_i);
    else ret = p->execute(c_i, fk_i);
    if (ret != 0) {
      cudaStreamSynchronize(stream);
      delete p;
      return ret;
    }
  }
  if (o.debug) cudaStreamSynchronize(stream);
  cudaError_t ce = cudaGetLastError();
  if (ce != cudaSuccess) {
    fprintf(stderr, "[b200nufft] CUDA error: %s\n", cudaGetErrorString(ce));
    cudaStreamSynchronize(stream);
    delete p;
    return B2N_ERR_CUDA_FAILURE;
  }
  cache_put(key, p, stream, cid);
  return warn;
}

int b2n_run(int type, int dim, int is_double, void *stream, double eps, int iflag, int64_t n_tot,
            int n_transf, int64_t n_j, const int64_t *n_k, const b2n_opts *opts, const void *src,
            const void *const *pts, const void *const *tgt, void *out) {
  return run_core(type, dim, is_double, (cudaStream_t)stream, eps, iflag, n_tot, n_transf, n_j, n_k, opts, src,
                  pts, tgt, out, nullptr);
}

// Host-buffer entry.  Device staging buffers are kept per process (grow-only).  Copies run on a
// private copy stream and are enqueued once the plan is known (ChunkCtx::enqueue):
//   * large single transforms (types 1, 2) arrive in chunks of points -- coordinates of chunk k,
//     then its strengths -- and the compute stream bin-sorts and spreads / interpolates chunk k
//     while chunk k+1 is on the wire; the uniform-grid stages run once (Plan::exec_phase).  For
//     type 2 each chunk's outputs return on a third stream while the next chunk is interpolated.
//     The timed region of bench.py's e2e is then the PCIe transfer plus the last chunk's work;
//   * everything else: coordinates first, the source array follows while the compute stream is
//     already bin-sorting; the result comes back on the compute stream.
namespace {
struct HostStage {
  std::mutex mu;
  void *buf[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};  // src, out, p0-2, t0-2
  size_t cap[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  cudaStream_t copy = nullptr, d2h = nullptr;
  cudaEvent_t pts_ready = nullptr, src_ready = nullptr, done = nullptr;
  static constexpr int MAXC = 8;
  cudaEvent_t pts_ev[MAXC], src_ev[MAXC], out_ev[MAXC];
  int device = -1;
};
HostStage g_stage;
int stage_grow(HostStage &h, int i, size_t bytes, cudaStream_t st) {
  if (bytes <= h.cap[i]) return 0;
  dev_free(h.buf[i], st);
  h.buf[i] = nullptr;
  h.cap[i] = 0;
  if (int e = dev_alloc(&h.buf[i], bytes, st)) return e;
  h.cap[i] = bytes;
  return 0;
}
struct HostJob {
  HostStage *h;
  int type, dim;
  int64_t n_tot, n_j;
  size_t rs, cs, src_b, pt_b, tg_b;
  const void *src;
  const void *const *pts;
  const void *const *tgt;
  cudaStream_t st;
};
// ChunkCtx::enqueue: all host-to-device copies of one call, in the order the compute stream
// consumes them
int host_enqueue(ChunkCtx *cc, PlanBase *plan) {
  HostJob &j = *(HostJob *)cc->user;
  HostStage &h = *j.h;
  int n = 0;
  int64_t ch = 0;
  const bool chunked = j.n_tot == 1 && j.type != 3 && plan->can_chunk(j.n_j, &n, &ch) && n <= HostStage::MAXC;
  if (!chunked) {
    cc->nchunk = 0;
    for (int d = 0; d < j.dim; d++) {
      B2N_CUDA_OK(cudaMemcpyAsync(h.buf[2 + d], j.pts[d], j.pt_b, cudaMemcpyHostToDevice, h.copy));
      if (j.type == 3) B2N_CUDA_OK(cudaMemcpyAsync(h.buf[5 + d], j.tgt[d], j.tg_b, cudaMemcpyHostToDevice, h.copy));
    }
    B2N_CUDA_OK(cudaEventRecord(h.pts_ready, h.copy));
    B2N_CUDA_OK(cudaMemcpyAsync(h.buf[0], j.src, j.src_b, cudaMemcpyHostToDevice, h.copy));
    B2N_CUDA_OK(cudaEventRecord(h.src_ready, h.copy));
    B2N_CUDA_OK(cudaStreamWaitEvent(j.st, h.pts_ready, 0));
    return 0;
  }
  cc->nchunk = n;
  cc->chunk = ch;
  cc->pts_ev = h.pts_ev;
  cc->src_ev = h.src_ev;
  cc->out_ev = h.out_ev;
  cc->d2h = h.d2h;
  if (j.type == 2) {  // the modes first: amplify + FFT overlap the first chunk's coordinates
    B2N_CUDA_OK(cudaMemcpyAsync(h.buf[0], j.src, j.src_b, cudaMemcpyHostToDevice, h.copy));
    B2N_CUDA_OK(cudaEventRecord(h.src_ev[0], h.copy));
  }
  for (int k = 0; k < n; k++) {
    const int64_t off = (int64_t)k * ch, m = std::min<int64_t>(ch, j.n_j - off);
    for (int d = 0; d < j.dim; d++)
      B2N_CUDA_OK(cudaMemcpyAsync((char *)h.buf[2 + d] + (size_t)off * j.rs, (const char *)j.pts[d] + (size_t)off * j.rs,
                                  (size_t)m * j.rs, cudaMemcpyHostToDevice, h.copy));
    B2N_CUDA_OK(cudaEventRecord(h.pts_ev[k], h.copy));
    if (j.type == 1) {
      B2N_CUDA_OK(cudaMemcpyAsync((char *)h.buf[0] + (size_t)off * j.cs, (const char *)j.src + (size_t)off * j.cs,
                                  (size_t)m * j.cs, cudaMemcpyHostToDevice, h.copy));
      B2N_CUDA_OK(cudaEventRecord(h.src_ev[k], h.copy));
    }
  }
  return 0;
}
}  // namespace

int b2n_run_host(int type, int dim, int is_double, double eps, int iflag, int64_t n_tot, int n_transf,
                 int64_t n_j, const int64_t *n_k, const b2n_opts *opts, const void *src,
                 const void *const *pts, const void *const *tgt, void *out) {
  if (dim < 1 || dim > 3) return B2N_ERR_DIM_NOTVALID;
  if (type < 1 || type > 3) return B2N_ERR_TYPE_NOTVALID;
  cudaStream_t st = opts && opts->gpu_stream ? (cudaStream_t)opts->gpu_stream : 0;
  int64_t n_k_total = 1;
  if (type != 3) for (int d = 0; d < dim; d++) n_k_total *= n_k[d];
  else n_k_total = n_k[0];
  const size_t rs = is_double ? 8 : 4, cs = 2 * rs;
  const size_t src_b = (size_t)n_tot * n_transf * (type == 2 ? n_k_total : n_j) * cs;
  const size_t out_b = (size_t)n_tot * n_transf * (type == 2 ? n_j : n_k_total) * cs;
  const size_t pt_b = (size_t)n_tot * n_j * rs, tg_b = (size_t)n_tot * n_k_total * rs;
  HostStage &h = g_stage;
  std::lock_guard<std::mutex> lk(h.mu);
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return B2N_ERR_CUDA_FAILURE;
  if (h.device != dev) {  // first use, or the caller switched device: drop the old staging area
    for (int i = 0; i < 8; i++) { if (h.buf[i]) cudaFree(h.buf[i]); h.buf[i] = nullptr; h.cap[i] = 0; }
    if (h.copy) {
      cudaStreamDestroy(h.copy); cudaStreamDestroy(h.d2h);
      cudaEventDestroy(h.pts_ready); cudaEventDestroy(h.src_ready); cudaEventDestroy(h.done);
      for (int k = 0; k < HostStage::MAXC; k++) { cudaEventDestroy(h.pts_ev[k]); cudaEventDestroy(h.src_ev[k]); cudaEventDestroy(h.out_ev[k]); }
    }
    B2N_CUDA_OK(cudaStreamCreateWithFlags(&h.copy, cudaStreamNonBlocking));
    B2N_CUDA_OK(cudaStreamCreateWithFlags(&h.d2h, cudaStreamNonBlocking));
    B2N_CUDA_OK(cudaEventCreateWithFlags(&h.pts_ready, cudaEventDisableTiming));
    B2N_CUDA_OK(cudaEventCreateWithFlags(&h.src_ready, cudaEventDisableTiming));
    B2N_CUDA_OK(cudaEventCreateWithFlags(&h.done, cudaEventDisableTiming));
    for (int k = 0; k < HostStage::MAXC; k++) {
      B2N_CUDA_OK(cudaEventCreateWithFlags(&h.pts_ev[k], cudaEventDisableTiming));
      B2N_CUDA_OK(cudaEventCreateWithFlags(&h.src_ev[k], cudaEventDisableTiming));
      B2N_CUDA_OK(cudaEventCreateWithFlags(&h.out_ev[k], cudaEventDisableTiming));
    }
    h.device = dev;
  }
  int rc = 0;
  if ((rc = stage_grow(h, 0, src_b, st)) || (rc = stage_grow(h, 1, out_b, st))) return rc;
  for (int d = 0; d < dim; d++) {
    if ((rc = stage_grow(h, 2 + d, pt_b, st))) return rc;
    if (type == 3 && (rc = stage_grow(h, 5 + d, tg_b, st))) return rc;
  }
  // the copy stream may only start once earlier work on `st` (previous call, allocations) is done
  B2N_CUDA_OK(cudaEventRecord(h.done, st));
  B2N_CUDA_OK(cudaStreamWaitEvent(h.copy, h.done, 0));
  HostJob job{&h, type, dim, n_tot, n_j, rs, cs, src_b, pt_b, tg_b, src, pts, tgt, st};
  ChunkCtx cc;
  cc.enqueue = host_enqueue;
  cc.user = &job;
  cc.host_out = out;
  void *d_p[3] = {h.buf[2], h.buf[3], h.buf[4]}, *d_t[3] = {h.buf[5], h.buf[6], h.buf[7]};
  rc = run_core(type, dim, is_double, st, eps, iflag, n_tot, n_transf, n_j, n_k, opts, h.buf[0], d_p,
                type == 3 ? d_t : nullptr, h.buf[1], h.src_ready, &cc);
  if (rc <= 1) {
    if (!(cc.nchunk >= 2 && type == 2))  // chunked type 2 has already sent its outputs home
      cudaMemcpyAsync(out, h.buf[1], out_b, cudaMemcpyDeviceToHost, st);
    if (cudaStreamSynchronize(st) != cudaSuccess) rc = B2N_ERR_CUDA_FAILURE;
    if (cudaStreamSynchronize(h.d2h) != cudaSuccess) rc = B2N_ERR_CUDA_FAILURE;
  } else {
    cudaStreamSynchronize(h.copy);
    cudaStreamSynchronize(h.d2h);
    cudaStreamSynchronize(st);
  }
  return rc;
}

int b2n_setup_spreader(double eps, double upsampfac, int kerevalmeth, int is_double, int *ns, double *beta) {
  return setup_spreader(eps, upsampfac, kerevalmeth, is_double != 0, ns, beta);
}
int64_t b2n_next235beven(int64_t n, int64_t b) { return next235beven(n, b); }
int64_t b2n_set_nf_type12(int64_t ms, double upsampfac, int ns) { return set_nf_type12(ms, upsampfac, ns); }
void b2n_fseries(int64_t nf, int ns, double beta, double *out) { fseries_host(nf, ns, beta, out); }
int b2n_horner_table(int ns, double beta, int is_double, double *coef) {
  return horner_fit(ns, beta, is_double != 0, coef);
}
void b2n_default_binsize(int dim, int ns, int is_double, int type, int *bin) {
  (void)type;
  bool ok = is_double ? choose_bins<double>(dim, ns, bin) : choose_bins<float>(dim, ns, bin);
  if (!ok) { bin[0] = bin[1] = bin[2] = 0; }
}

}  // extern "C"
