// plan.h -- plan state + internal stage entry points of libb200nufft.
#pragma once
#include <vector>

#include "common.cuh"
#include "hostmath.h"

namespace b2n {

struct StageTimer;

// One sorted point: folded coordinates + original index, 16 B (float) / 32 B (double), so the
// spread/interp kernels read a point with one (two) LDG.128 and the sort moves it as a unit.
template <typename T> struct PtRec;
template <> struct __align__(16) PtRec<float> {
  float x, y, z;
  int32_t idx;
};
template <> struct __align__(16) PtRec<double> {
  double x, y, z;
  int64_t idx;
};

// Geometry of one bin-sorted point set (the state cuFINUFFT keeps in idxnupts / binsize /
// binstartpts / subprob_to_bin -- V/include/cufinufft/types.h:30-101).
template <typename T> struct PointSet {
  int64_t M = 0;
  PtRec<T> *rec = nullptr;       // points in key order                               [M]
  PtRec<T> *tmp = nullptr;       // partition scratch (sort pass P1)                  [M]
  int32_t *idx = nullptr;        // sorted position -> original index; materialised only for
                                 // b2n_plan_sort_get                                  [M]
  int32_t *key_cnt = nullptr;    // key histogram, counted down to zero by pass P2    [K]
  int32_t *key_start = nullptr;  // exclusive scan of the key histogram               [K+1]
  int32_t *bin_start = nullptr;  // exclusive scan of the BIN histogram               [nbins+1]
  int32_t *sp_off = nullptr;     // exclusive scan of subproblems per bin             [nbins+1]
  int32_t *sp_bin = nullptr;     // subproblem -> bin                                 [sp_cap]
  int32_t *bucket_cur = nullptr; // per-bucket write cursors of pass P1               [256]
  int64_t sp_cap = 0;            // host-known upper bound on #subproblems
  int nsub = 1;                  // sub-keys per bin (z cells per bin for the SWR kernels)
  bool sorted = false;
  // setpts cache (sort.cu): device {accumulator, signature of the sorted set, skip flag}
  unsigned long long *sig = nullptr;
  unsigned long long sig_salt = 0;  // M + sort geometry of the sorted set
  bool sig_ok = false;              // the last binsort_points completed with the cache on
  // two-pass sort: the overflow flag of the last attempt, copied back asynchronously (pinned; never
  // waited for) -- a point set that overflowed makes the next calls skip the attempt (fast_hold)
  int *ovf_host = nullptr;
  int fast_hold = 0;
  int64_t cap_M = 0, cap_tmp = 0, cap_idx = 0, cap_keys = 0, cap_bins = 0, cap_sp = 0;
};

struct PlanBase {
  virtual ~PlanBase() {}
  virtual int setpts(int64_t M, const void *x, const void *y, const void *z, int64_t N,
                     const void *s, const void *t, const void *u) = 0;
  virtual int execute(void *c, void *fk) = 0;
  // one transform in pieces, for point sets that arrive in chunks (b2n_run_host): PH_BEGIN =
  // what precedes the non-uniform stage (type 1: clear the fine grid; type 2: amplify + FFT),
  // PH_BODY = spread / interpolate the CURRENT point set (c = that chunk's strengths / outputs),
  // PH_END = what follows it (type 1: FFT + deconvolve).  Types 1 and 2, ntransf <= batch.
  enum { PH_BEGIN = 1, PH_BODY = 2, PH_END = 4 };
  virtual int exec_phase(int phase, void *c, void *fk) = 0;
  virtual bool can_chunk(int64_t M, int *nchunk, int64_t *chunk) const = 0;
  // What precedes the non-uniform stage (PH_BEGIN: type 1 clears the fine grid, type 2 amplifies
  // and FFTs the modes) does not depend on the points: b2n_run starts it on the plan's side stream
  // BEFORE the bin-sort and joins it afterwards, so the two overlap (the sort is bound by atomics
  // and latency, amplify + FFT by DRAM bandwidth).  overlap_begin returns false if the plan cannot
  // (type 3, more transforms than a batch, spread/interp only, debug timers on).
  virtual bool overlap_begin(void *c, void *fk, cudaEvent_t wait_first, int *err) = 0;
  virtual void overlap_join() = 0;
  virtual void info(b2n_plan_info *out) = 0;
  virtual int sort_get(const int32_t **idx, const int32_t **bin_start, int64_t *nbins) = 0;
  virtual void set_stream(cudaStream_t s) = 0;
  virtual cudaStream_t get_stream() const = 0;
  bool is_double = false;
  int plan_warning = 0;  // what makeplan returned (1 = eps below machine precision): repeated on cache hits
  double timings[7] = {0, 0, 0, 0, 0, 0, 0};
};

template <typename T> struct Plan : PlanBase {
  int type = 0, dim = 0, iflag = 1, ntransf = 1, batch = 1;
  double eps = 0;
  b2n_opts opts;
  int method = 0;  // 1 = GM kernels, 2 = tile kernels, 3 = register kernels (3-D sliding window, 2-D tile)
  int ns = 0, ncoef = 0;
  double beta = 0, sigma = 2.0;
  int warn = 0;
  int64_t ms[3] = {1, 1, 1};
  int64_t nmodes = 1;
  int64_t nf[3] = {1, 1, 1};
  int64_t nftot = 1;
  int bin[3] = {1, 1, 1};
  int nbin[3] = {1, 1, 1};
  int64_t nbins = 1;
  int maxsub = 1024;
  // geometry alternatives: the tile/GM choice made at plan time, and whether the sliding-window
  // register kernels may take over once the point count is known (set_geometry, at setpts)
  int base_method = 0, base_bin[3] = {1, 1, 1}, base_maxsub = 1024;
  bool swr_ok = false;
  double sort_fill = 1.0;  // fraction of the slowest axis the points are known to occupy (binsort_points: bucket capacity)
  bool stacked2 = false;  // 2-D type 1 with stacked transforms: narrow-window bins (rt2_kernels.cuh, Rt2sCfg)
  cpx<T> *cpack = nullptr;  // point-major strengths of the stacked transforms of one spread launch [M][8 * passes]
  int64_t cap_cpack = 0;
  cpx<T> *fw_stack = nullptr;  // stacked 2-D type 1: fine grids of a whole GROUP of batches (exec1_stacked)
  int64_t cap_fw_stack = 0;
  HornerTable<T> tab;
  cudaStream_t stream = 0;
  cudaStream_t side = nullptr;             // overlap_begin / overlap_join
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

  T *fwker[3] = {nullptr, nullptr, nullptr};  // kernel Fourier series, nf_d/2+1 each
  cpx<T> *fw = nullptr;                       // fine grid(s): batch * nftot
  cufftHandle fft = 0;
  bool has_fft = false;
  // 3-D types 1/2: only N3 of the nf3 z-planes carry modes, so the (y, x) transforms of the other
  // planes are never needed: one strided 1-D plan along z over the whole grid + batched 2-D plans
  // on the two slabs of kept planes (2 passes over the data instead of 3)
  cufftHandle fft_z = 0, fft_xy[2] = {0, 0};
  int64_t slab_lo[2] = {0, 0}, slab_n[2] = {0, 0};
  bool pruned = false;

  PointSet<T> pts;

  // ---- type 3 (V/include/cufinufft/types.h:84-100 type3_params + prephase/deconv) ----
  int64_t N3 = 0;
  double t3X[3] = {0, 0, 0}, t3C[3] = {0, 0, 0}, t3S[3] = {0, 0, 0}, t3D[3] = {0, 0, 0};
  double t3h[3] = {1, 1, 1}, t3gam[3] = {1, 1, 1};
  T *xp[3] = {nullptr, nullptr, nullptr};  // rescaled sources x' [M]
  T *sp[3] = {nullptr, nullptr, nullptr};  // rescaled targets s' [N]
  cpx<T> *prephase = nullptr;              // [M]
  cpx<T> *deconv = nullptr;                // [N]
  Plan<T> *inner = nullptr;                // inner type-2 plan (impl.h:795-812)
  int64_t cap_xp = 0, cap_sp3 = 0, cap_fw = 0;

  ~Plan() override;
  int init(int type, int dim, const int64_t *n_modes, int iflag, int ntransf, double eps,
           const b2n_opts *opts);
  int alloc_grid();
  void set_geometry(int64_t M);
  int setpts(int64_t M, const void *x, const void *y, const void *z, int64_t N, const void *s,
             const void *t, const void *u) override;
  int setpts12(int64_t M, const T *x, const T *y, const T *z);
  int setpts3(int64_t M, const T *x, const T *y, const T *z, int64_t N, const T *s, const T *t,
              const T *u);
  int execute(void *c, void *fk) override;
  int exec_phase(int phase, void *c, void *fk) override;
  bool can_chunk(int64_t M, int *nchunk, int64_t *chunk) const override;
  bool overlap_begin(void *c, void *fk, cudaEvent_t wait_first, int *err) override;
  void overlap_join() override;
  int spread(const cpx<T> *c, const cpx<T> *prescale, cpx<T> *grid, int ntr);
  int interp(cpx<T> *c, const cpx<T> *postscale, const cpx<T> *grid, int ntr);
  int exec1(cpx<T> *c, cpx<T> *fk);
  int exec1_stacked(cpx<T> *c, cpx<T> *fk, int group);
  int exec2(cpx<T> *c, cpx<T> *fk, const cpx<T> *postscale);
  int exec3(cpx<T> *c, cpx<T> *fk);
  void info(b2n_plan_info *out) override;
  int sort_get(const int32_t **idx, const int32_t **bin_start, int64_t *nbins) override;
  void set_stream(cudaStream_t s) override;
  cudaStream_t get_stream() const override { return stream; }
};

// ---- stage launchers (one per .cu file) -------------------------------------------------------
template <typename T>
int binsort_points(Plan<T> &p, int64_t M, const T *x, const T *y, const T *z);
template <typename T> int materialise_idx(Plan<T> &p);

// strengths c (gathered through pts.idx, optionally multiplied by prescale[idx]) -> fw (+=)
template <typename T>
int spread_tile(Plan<T> &p, const cpx<T> *c, const cpx<T> *prescale, cpx<T> *fw, int ntr);
template <typename T>
int spread_gm(Plan<T> &p, const cpx<T> *c, const cpx<T> *prescale, cpx<T> *fw, int ntr);
// fw -> c[idx] (optionally multiplied by postscale[idx])
template <typename T>
int interp_tile(Plan<T> &p, cpx<T> *c, const cpx<T> *postscale, const cpx<T> *fw, int ntr);
template <typename T>
int interp_gm(Plan<T> &p, cpx<T> *c, const cpx<T> *postscale, const cpx<T> *fw, int ntr);

// register kernels (swr.cu): sliding window (3-D) / register tile (2-D), float, ns <= 8, bins from swr_bins()
void swr_bins(int dim, int ns, int *bin, bool stacked2 = false);
int spread_swr(Plan<float> &p, const float2 *c, const float2 *prescale, float2 *fw, int ntr);
int interp_swr(Plan<float> &p, float2 *c, const float2 *postscale, const float2 *fw, int ntr);

// smem bytes the tile kernels need for (dim, ns, bins); 0 if the combination is unsupported
template <typename T> size_t tile_smem_bytes(int dim, int ns, const int *bin);

template <typename T>
int deconvolve(Plan<T> &p, const cpx<T> *fw, cpx<T> *fk, int ntr);   // type 1 step 3
template <typename T>
int amplify(Plan<T> &p, cpx<T> *fw, const cpx<T> *fk, int ntr);      // type 2 step 1 (+zero pad)
template <typename T> int compute_fseries(Plan<T> &p);

template <typename T>
int t3_minmax(cudaStream_t st, int dim, int64_t M, const T *const *x, int64_t N,
              const T *const *s, double *lohi /* [12]: per dim lo,hi of x then of s */);
template <typename T> int t3_prepare(Plan<T> &p, const T *const *x, const T *const *s);

int dev_alloc(void **p, size_t bytes, cudaStream_t st);   // from the library's private pool (sort.cu)
void pool_usage(size_t *reserved, size_t *used);
void pool_trim(size_t keep_bytes);
void dev_free(void *p, cudaStream_t st);
template <typename U> inline int dev_alloc_t(U **p, size_t count, cudaStream_t st) {
  return dev_alloc((void **)p, count * sizeof(U), st);
}
int exclusive_scan_i32(const int32_t *in, int32_t *out, int64_t n, cudaStream_t st, const int *skip = nullptr);
bool setpts_cache_enabled();  // api.cu: b2n_set_setpts_cache / B2N_SETPTS_CACHE

}  // namespace b2n
