/*
 * b200nufft.h -- C ABI of libb200nufft.so: the B200-native (sm_100a) NUFFT backend that replaces
 * jax-finufft's GPU path (lib/jax_finufft_gpu.cc + lib/kernels.cc.cu + lib/cufinufft_wrapper.*
 * and, below them, cuFINUFFT).  Plain pointers and sizes only -- no torch / XLA / C++ types.
 *
 * All data pointers are DEVICE pointers unless the function name ends in `_host`.
 * All functions return the FINUFFT integer error convention of the reference
 * (vendor/finufft/include/finufft_errors.h:6-32): 0 ok, 1 = warning "eps too small" (NOT an
 * error, lib/kernels.cc.cu:52), >1 error.  Nothing throws across this boundary.
 *
 * Paths below are relative to the reference tree; V/ = vendor/finufft/.
 */
#ifndef B200NUFFT_H
#define B200NUFFT_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Error codes: identical values to V/include/finufft_errors.h:6-32 */
enum {
  B2N_OK = 0,
  B2N_WARN_EPS_TOO_SMALL = 1,
  B2N_ERR_MAXNALLOC = 2,
  B2N_ERR_UPSAMPFAC_TOO_SMALL = 7,
  B2N_ERR_HORNER_WRONG_BETA = 8,
  B2N_ERR_NTRANS_NOTVALID = 9,
  B2N_ERR_TYPE_NOTVALID = 10,
  B2N_ERR_ALLOC = 11,
  B2N_ERR_DIM_NOTVALID = 12,
  B2N_ERR_NDATA_NOTVALID = 14,
  B2N_ERR_CUDA_FAILURE = 15,
  B2N_ERR_PLAN_NOTVALID = 16,
  B2N_ERR_METHOD_NOTVALID = 17,
  B2N_ERR_BINSIZE_NOTVALID = 18,
  B2N_ERR_INSUFFICIENT_SHMEM = 19,
  B2N_ERR_NUM_NU_PTS_INVALID = 20,
  B2N_ERR_INVALID_ARGUMENT = 21
};

/* Replaces `cufinufft_opts` (V/include/cufinufft_opts.h).  The first seven fields are exactly
 * the attributes that cross jax-finufft's FFI boundary (lib/jax_finufft_gpu.cc:28-60,
 * src/jax_finufft/lowering.py:157-174); the rest exist in the reference struct and are kept for
 * plan-level users (tests, type-3 internals). */
typedef struct b2n_opts {
  int modeord;            /* 0: modes from -N/2 up (CMCL); 1: FFT order                        */
  double upsampfac;       /* sigma; 0 = auto (2.0; 1.25 for type 3 when eps>=1e-9, impl.h:151) */
  int gpu_method;         /* 0 auto, 1 GM(-sort) global atomics/gathers, 2 SM tile, 3 OD -> tile */
  int gpu_sort;           /* 1: bin-sort the points (required by the tile kernels)             */
  int gpu_kerevalmeth;    /* 1: piecewise-polynomial (Horner) ES kernel; 0: exp(sqrt) direct    */
  int gpu_maxbatchsize;   /* 0 = min(ntransf, 8) as V/include/cufinufft/impl.h:123-127         */
  int debug;              /* 1: print plan parameters and per-stage timings                    */
  int gpu_binsizex, gpu_binsizey, gpu_binsizez; /* 0 = backend default (B200-tuned)            */
  int gpu_maxsubprobsize; /* 0 = backend default                                                */
  int gpu_spreadinterponly; /* 1: no FFT/deconvolve: fk IS the fine grid (impl.h:115-117)      */
  int gpu_device_id;      /* informational; the current device at call time is used            */
  void *gpu_stream;       /* cudaStream_t all work is enqueued on                              */
} b2n_opts;

typedef struct b2n_plan_s *b2n_plan;

/* Plan parameters, for tests and debug printing (what `debug=1` prints in the reference,
 * V/src/cuda/common.cu:550-566, V/include/cufinufft/impl.h:169-173). */
typedef struct b2n_plan_info {
  int type, dim, is_double, ns, method, ntransf, batchsize, ncoef;
  double beta, upsampfac;
  int64_t nf[3], ms[3];
  int binsize[3], nbins[3];
  int64_t M, N;
  int64_t t3_nf_inner[3];
  double t3_X[3], t3_C[3], t3_S[3], t3_D[3], t3_h[3], t3_gam[3];
} b2n_plan_info;

/* replaces cufinufft_default_opts (V/src/cuda/cufinufft.cu:120-152) */
void b2n_default_opts(b2n_opts *opts);

/* replaces cufinufft{,f}_makeplan (V/include/cufinufft.h:19-24, V/src/cuda/cufinufft.cu:32-63,
 * V/include/cufinufft/impl.h:52-334).  n_modes has 3 entries (x fastest); unused dims ignored. */
int b2n_makeplan(int type, int dim, const int64_t *n_modes, int iflag, int ntransf, double eps,
                 int is_double, b2n_plan *plan, const b2n_opts *opts);

/* replaces cufinufft{,f}_setpts (V/include/cufinufft.h:26-31, impl.h:337-823).
 * x,y,z: M source points (float* or double* per the plan precision); s,t,u: N targets, type 3. */
int b2n_setpts(b2n_plan plan, int64_t M, const void *x, const void *y, const void *z, int64_t N,
               const void *s, const void *t, const void *u);

/* replaces cufinufft{,f}_execute (V/include/cufinufft.h:33-36, impl.h:826-870).
 * c: [ntransf][M] complex; fk: [ntransf][ms*mt*mu] (types 1,2) or [ntransf][N] (type 3). */
int b2n_execute(b2n_plan plan, void *c, void *fk);

/* replaces cufinufft{,f}_destroy (V/include/cufinufft.h:38-39, impl.h:872-900) */
int b2n_destroy(b2n_plan plan);

int b2n_plan_info_get(b2n_plan plan, b2n_plan_info *info);

/* Sorted-permutation introspection (the contract of V/src/cuda/3d/spreadinterp3d.cuh:28-84):
 * idx[M] (idxnupts), bin_start[nbins+1] (exclusive scan of the bin histogram, x-fastest bins).
 * Device pointers owned by the plan, valid until the next setpts/destroy. */
int b2n_plan_sort_get(b2n_plan plan, const int32_t **idx, const int32_t **bin_start,
                      int64_t *nbins);
/* Copies the same two arrays into caller-owned DEVICE buffers (idx_out[M], bin_start_out[nbins+1])
 * on the plan's stream and synchronises it. */
int b2n_plan_sort_copy(b2n_plan plan, int32_t *idx_out, int32_t *bin_start_out);

/* replaces run_nufft<ndim,T,type> (lib/kernels.cc.cu:25-92) -- the one call the XLA-FFI shim
 * makes per custom call: plan once, loop n_tot point sets {setpts, execute}, no retained
 * pointers.  Layouts as the FFI delivers them (src/jax_finufft/lowering.py:96-122):
 *   src  (n_tot, n_transf, n_j) for types 1,3; (n_tot, n_transf, n_k3?, n_k2?, n_k1) for type 2
 *   pts  dim arrays (n_tot, n_j), x (fastest grid dim) first
 *   tgt  dim arrays (n_tot, n_k[0]) for type 3, else NULLs
 *   out  (n_tot, n_transf, n_k...) types 1; (n_tot, n_transf, n_j) type 2; (.., n_k[0]) type 3
 * Unlike the reference it does not block the host (no trailing cudaStreamSynchronize) unless
 * opts->debug is set; plans and workspaces come from a per-process cache. */
int b2n_run(int type, int dim, int is_double, void *stream, double eps, int iflag, int64_t n_tot,
            int n_transf, int64_t n_j, const int64_t *n_k, const b2n_opts *opts, const void *src,
            const void *const *pts, const void *const *tgt, void *out);

/* Same call with HOST buffers (pinned or pageable): stages inputs to the device, runs, copies
 * the result back and synchronises.  This is what bench.py's `e2e` number times. */
int b2n_run_host(int type, int dim, int is_double, double eps, int iflag, int64_t n_tot,
                 int n_transf, int64_t n_j, const int64_t *n_k, const b2n_opts *opts,
                 const void *src, const void *const *pts, const void *const *tgt, void *out);

/* ---- the custom-call boundary, flattened to C -------------------------------------------------
 * One XLA custom call of jax-finufft = one b2n_ffi_call.  `target` is the custom-call target
 * name the reference registers (lib/jax_finufft_gpu.cc:356-422): "nufft{1,2,3}d{1,2,3}" with a
 * trailing "f" for single precision.  `attrs` carries the typed FFI attributes in the order of
 * lib/jax_finufft_gpu.cc:28-60 (src/jax_finufft/lowering.py:157-174); eps is a float attribute
 * for the ...f targets upstream and is widened here.  `operands` are the device pointers of the
 * call's arguments in FFI order: source, points x dim, then (type 3) targets x dim -- points
 * already reversed so that operands[1] is the fastest grid axis (lowering.py:104-105).
 * `result` is the output buffer.  Returns the FINUFFT code; 0 and 1 map to ffi::Error::Success()
 * (lib/kernels.cc.cu:52), anything else to ffi::Error::Internal(b2n_strerror(code)). */
typedef struct b2n_ffi_attrs {
  double eps;
  int64_t iflag, n_tot, n_transf, n_j, n_k_1, n_k_2, n_k_3, modeord;
  double upsampfac;
  int64_t gpu_method, gpu_sort, gpu_kerevalmeth, gpu_maxbatchsize, debug;
} b2n_ffi_attrs;

int b2n_ffi_call(const char *target, void *stream, const b2n_ffi_attrs *attrs,
                 const void *const *operands, int n_operands, void *result);
/* Number of operands the target takes (1 + dim, or 1 + 2*dim for type 3); -1: unknown target. */
int b2n_ffi_arity(const char *target);
/* The 18 target names, NULL-terminated (what registrations() returns as keys). */
const char *const *b2n_ffi_targets(void);
/* Message for an error code, in the wording of lib/kernels.cc.cu:54,72,78,88. */
const char *b2n_strerror(int code);

/* Drop every cached plan / workspace of the calling process and return the memory of the
 * library's private stream-ordered pool to the driver (tests, memory pressure).  The reference
 * holds no state between calls (it builds and destroys its plan per call, lib/kernels.cc.cu:49-51,
 * 84); these three entry points are what a host framework uses to bound ours. */
void b2n_cache_clear(void);
/* Byte bound of the plan cache: after a call parks its plan, the oldest parked plans are dropped
 * until the pool's used bytes fit.  Default: a quarter of the device (environment
 * B2N_CACHE_BYTES); bytes < 0 restores the default.  Returns the previous limit. */
long long b2n_set_cache_limit(long long bytes);
/* Bytes the library's pool on the current device holds from the driver / has handed out. */
void b2n_cache_bytes(unsigned long long *reserved, unsigned long long *used);

/* Setpts cache (the reference re-sorts on every custom call,
 * lib/kernels.cc.cu:49-51,64).  When on, setpts folds the coordinate arrays into a 64-bit
 * signature and compares it on the device with that of the point set the (cached) plan already
 * holds sorted; on a match the bin-sort kernels return at once -- no host round trip, the call
 * stays stream-ordered.  Off by default (also: environment B2N_SETPTS_CACHE=1).  Returns the
 * previous setting. */
int b2n_set_setpts_cache(int on);

/* Spatial multi-GPU split of a 3-D type 1 (jax_finufft_b200/parallel.py, combine="slab"; the
 * reference shards above its custom call, tests/sharding_test.py:119-194): groups this rank's M
 * points by the rank that owns their fine-grid plane along the slowest axis (p0; rank r owns planes
 * [r*nf0/world, (r+1)*nf0/world)) and re-bases that coordinate to the owner's local grid of
 * nf0/world + 2*halo planes.  o0 (re-based p0), o1, o2: real[M]; oc: complex[M] -- each grouped
 * by owner, same order; counts2: 2*world uint64 on the device, [0, world) = points per owner on
 * return.  world <= 16. */
int b2n_slab_partition(int is_double, void *stream, int64_t M, const void *p0, const void *p1,
                       const void *p2, const void *c, int64_t nf0, int world, int halo, void *o0,
                       void *o1, void *o2, void *oc, void *counts2);

/* The uniform-grid stages of a type 1 whose fine grid is distributed over the GPUs in z-slabs
 * (parallel.py: slab_pencil_fft; SURVEY.md 8(e) native path; the reference's equivalent is the
 * replicated FFT + psum of tests/sharding_test.py:163-165).  All work is enqueued on `stream`;
 * cuFFT plans and kernel series are cached per geometry.
 *  b2n_slab_fft_xy: `slab` = this rank's nzl planes (nzl, nf2, nf1), transformed IN PLACE over
 *    (y, x); `send` (nzl * n2 * n1 complex) receives the n1 x n2 central modes divided by the x / y
 *    kernel series (index maps: V/src/cuda/deconvolve_wrapper.cu:76-118), grouped by the rank that
 *    owns their y range (contiguous blocks, the first n2 % world ranks one row longer):
 *    send[dst][z][y - lo(dst)][x] -- the operand of one all_to_all.
 *  b2n_slab_fft_z: `pencil` = the received block (nf3, plane = n2_local * n1), z-major,
 *    transformed IN PLACE along z; out (n3, n2_local, n1) = its n3 central modes / z series. */
int b2n_slab_fft_xy(int is_double, void *stream, void *slab, int64_t nzl, int64_t nf2, int64_t nf1,
                    int64_t n2, int64_t n1, int world, int iflag, int modeord, int ns, double beta,
                    void *send);
int b2n_slab_fft_z(int is_double, void *stream, void *pencil, int64_t nf3, int64_t n3, int64_t plane,
                   int iflag, int modeord, int ns, double beta, void *out);

/* The per-point stacks and reductions of the JVP / VJP rules (ref src/jax_finufft/ops.py:238-273,
 * 280-314: jnp.stack([c, dx*c, dy*c, dz*c], axis=2) and sum(conj(c) * h_d)), one fused pass each.
 *  b2n_stack_scaled: src complex [n_tot][n_transf][n]; scales[k], k < n_scale <= 4: real
 *    [n_tot][n] shared by the n_transf transforms, or NULL (factor 1);
 *    out[i][t][k][j] = scales[k][i][j] * src[i][t][j]   -- the operand of the stacked transform.
 *  b2n_grad_points: c complex [n_tot][n_transf][n]; h complex [n_tot][n_transf][n_comp][n];
 *    out[k][i][j] = sign * sum_t Im(conj(c[i][t][j]) * h[i][t][first + k][j])   (mode 0; mode 1: Re),
 *    k < count <= 4 -- the cotangents of the point coordinates. */
int b2n_stack_scaled(int is_double, void *stream, int64_t n_tot, int n_transf, int64_t n, int n_scale,
                     const void *src, const void *const *scales, void *out);
int b2n_grad_points(int is_double, void *stream, int64_t n_tot, int n_transf, int64_t n, int n_comp,
                    int first, int count, int mode, double sign, const void *c, const void *h, void *out);

/* Per-stage device timings (ms) of the most recent b2n_execute / b2n_setpts on this plan when
 * opts.debug != 0: [0] sort, [1] spread, [2] fft, [3] deconvolve/amplify, [4] interp,
 * [5] type-3 pre/post, [6] memset. */
int b2n_plan_timings(b2n_plan plan, double *ms7);

/* Host-side plan arithmetic, exported so `-m "not gpu"` tests can check it against the oracle
 * (V/src/cuda/spreadinterp.cpp:48-90, V/src/cuda/common.cu:166-209, V/src/common/utils.cpp). */
int b2n_setup_spreader(double eps, double upsampfac, int kerevalmeth, int is_double, int *ns,
                       double *beta);
int64_t b2n_next235beven(int64_t n, int64_t b);
int64_t b2n_set_nf_type12(int64_t ms, double upsampfac, int ns);
void b2n_fseries(int64_t nf, int ns, double beta, double *fwkerhalf /* nf/2+1 */);
/* piecewise-polynomial table of the ES kernel: coef[k*16 + j], k=0..ncoef-1 (highest power
 * first), j = interval 0..ns-1, variable z = 2*x1 + ns - 1 in [-1,1]; returns ncoef. */
int b2n_horner_table(int ns, double beta, int is_double, double *coef /* 24*16 */);
void b2n_default_binsize(int dim, int ns, int is_double, int type, int *binsize3);
const char *b2n_version(void);
/* Kernels of this library launched so far by the calling process (bench.py: gpu_launches). */
unsigned long long b2n_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* B200NUFFT_H */
