// ffi_core.cpp -- the custom-call boundary in plain C: target-name decoding, attribute -> opts
// mapping and operand routing for the 18 targets jax-finufft registers
// (lib/jax_finufft_gpu.cc:356-422).  The reference spells these out as 18 wrapper functions
// plus 12 binding builders; here one table-free decoder serves all of them, and the XLA-FFI
// shim (xla_ffi_shim.cc) is a single variadic handler on top of it.  Host code only.
#include <cstdio>
#include <cstring>

#include "../../include/b200nufft.h"

namespace {

struct Target {
  int dim, type, is_double;
};

// "nufft<dim>d<type>[f]"
bool decode(const char *name, Target *t) {
  if (!name || std::strncmp(name, "nufft", 5) != 0) return false;
  const size_t n = std::strlen(name);
  if (n != 8 && n != 9) return false;
  if (name[5] < '1' || name[5] > '3' || name[6] != 'd' || name[7] < '1' || name[7] > '3') return false;
  if (n == 9 && name[8] != 'f') return false;
  t->dim = name[5] - '0';
  t->type = name[7] - '0';
  t->is_double = n == 8;
  return true;
}

const char *const kTargets[] = {"nufft1d1f", "nufft1d1", "nufft2d1f", "nufft2d1", "nufft3d1f", "nufft3d1",
                                "nufft1d2f", "nufft1d2", "nufft2d2f", "nufft2d2", "nufft3d2f", "nufft3d2",
                                "nufft1d3f", "nufft1d3", "nufft2d3f", "nufft2d3", "nufft3d3f", "nufft3d3",
                                nullptr};

}  // namespace

extern "C" {

const char *const *b2n_ffi_targets(void) { return kTargets; }

int b2n_ffi_arity(const char *target) {
  Target t;
  if (!decode(target, &t)) return -1;
  return 1 + (t.type == 3 ? 2 : 1) * t.dim;
}

const char *b2n_strerror(int code) {
  switch (code) {
    case B2N_OK: return "success";
    case B2N_WARN_EPS_TOO_SMALL: return "warning: eps too small for this precision (not an error)";
    case B2N_ERR_MAXNALLOC: return "b200nufft: fine grid too large";
    case B2N_ERR_UPSAMPFAC_TOO_SMALL: return "b200nufft makeplan failed: upsampfac too small";
    case B2N_ERR_HORNER_WRONG_BETA: return "b200nufft makeplan failed: gpu_kerevalmeth=1 needs upsampfac 2 or 1.25";
    case B2N_ERR_NTRANS_NOTVALID: return "b200nufft makeplan failed: n_transf must be >= 1";
    case B2N_ERR_TYPE_NOTVALID: return "b200nufft makeplan failed: type must be 1, 2 or 3";
    case B2N_ERR_ALLOC: return "b200nufft: device allocation failed";
    case B2N_ERR_DIM_NOTVALID: return "b200nufft makeplan failed: dim must be 1, 2 or 3";
    case B2N_ERR_NDATA_NOTVALID: return "b200nufft makeplan failed: invalid number of modes or points";
    case B2N_ERR_CUDA_FAILURE: return "b200nufft: CUDA error";
    case B2N_ERR_PLAN_NOTVALID: return "b200nufft: invalid plan";
    case B2N_ERR_METHOD_NOTVALID: return "b200nufft makeplan failed: invalid gpu_method";
    case B2N_ERR_BINSIZE_NOTVALID: return "b200nufft makeplan failed: invalid bin size";
    case B2N_ERR_INSUFFICIENT_SHMEM: return "b200nufft makeplan failed: bins do not fit in shared memory";
    case B2N_ERR_NUM_NU_PTS_INVALID: return "b200nufft setpts failed: invalid number of nonuniform points";
    case B2N_ERR_INVALID_ARGUMENT: return "b200nufft: invalid argument";
    default: return "b200nufft: unknown error code";
  }
}

int b2n_ffi_call(const char *target, void *stream, const b2n_ffi_attrs *a, const void *const *operands,
                 int n_operands, void *result) {
  Target t;
  if (!decode(target, &t)) {
    std::fprintf(stderr, "[b200nufft] unknown custom-call target '%s'\n", target ? target : "(null)");
    return B2N_ERR_INVALID_ARGUMENT;
  }
  if (!a || !operands || n_operands != b2n_ffi_arity(target)) return B2N_ERR_INVALID_ARGUMENT;
  if (a->n_tot < 0 || a->n_j < 0 || a->n_transf < 1 || a->n_transf > 0x7fffffffLL) return B2N_ERR_NTRANS_NOTVALID;
  // A buffer may be NULL exactly when it is empty (torch's data_ptr() of a zero-size tensor, XLA's
  // empty buffers): nufft1 / nufft3 with n_j == 0 must return zeros, not an error.
  {
    int64_t n_modes = 1;
    for (int d = 0; d < t.dim; d++) n_modes *= (d == 0 ? a->n_k_1 : d == 1 ? a->n_k_2 : a->n_k_3);
    const int64_t n_uni = t.type == 3 ? a->n_k_1 : n_modes;            // modes, or type-3 targets
    const int64_t n_src = t.type == 2 ? n_uni : a->n_j, n_out = t.type == 2 ? a->n_j : n_uni;
    if (!operands[0] && a->n_tot * n_src > 0) return B2N_ERR_INVALID_ARGUMENT;
    if (!result && a->n_tot * n_out > 0) return B2N_ERR_INVALID_ARGUMENT;
    for (int d = 0; d < t.dim; d++) {
      if (!operands[1 + d] && a->n_tot * a->n_j > 0) return B2N_ERR_INVALID_ARGUMENT;
      if (t.type == 3 && !operands[1 + t.dim + d] && a->n_tot * a->n_k_1 > 0) return B2N_ERR_INVALID_ARGUMENT;
    }
  }

  // build_opts<T> of the reference (lib/kernels.cc.cu:98-113): start from the defaults, then the
  // seven attributes that cross the boundary
  b2n_opts o;
  b2n_default_opts(&o);
  o.modeord = (int)a->modeord;
  o.upsampfac = a->upsampfac;
  o.gpu_method = (int)a->gpu_method;
  o.gpu_sort = (int)a->gpu_sort;
  o.gpu_kerevalmeth = (int)a->gpu_kerevalmeth;
  o.gpu_maxbatchsize = (int)a->gpu_maxbatchsize;
  o.debug = (int)a->debug;

  const int64_t n_k[3] = {a->n_k_1, a->n_k_2, a->n_k_3};
  const void *pts[3] = {nullptr, nullptr, nullptr}, *tgt[3] = {nullptr, nullptr, nullptr};
  for (int d = 0; d < t.dim; d++) {
    pts[d] = operands[1 + d];
    if (t.type == 3) tgt[d] = operands[1 + t.dim + d];
  }
  return b2n_run(t.type, t.dim, t.is_double, stream, a->eps, (int)a->iflag, a->n_tot, (int)a->n_transf, a->n_j,
                 n_k, &o, operands[0], pts, tgt, result);
}

}  // extern "C"
