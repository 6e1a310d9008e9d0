"""ctypes binding of ``libb200nufft.so`` (the C ABI declared in ``include/b200nufft.h``).

The library is built in-tree by ``make -C jax_finufft_b200/csrc`` (see ``build()``).  There is
no fallback of any kind: if the shared object is missing or a call is made without a CUDA
device, an exception is raised.
"""

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
# B2N_LIB: load another build of the same library (kernel A/B experiments; never a different backend)
LIB_PATH = os.environ.get("B2N_LIB") or os.path.join(_HERE, "libb200nufft.so")
_lib = None


class B2nOpts(C.Structure):
    """``b2n_opts`` (include/b200nufft.h) -- replaces ``cufinufft_opts``."""

    _fields_ = [
        ("modeord", C.c_int),
        ("upsampfac", C.c_double),
        ("gpu_method", C.c_int),
        ("gpu_sort", C.c_int),
        ("gpu_kerevalmeth", C.c_int),
        ("gpu_maxbatchsize", C.c_int),
        ("debug", C.c_int),
        ("gpu_binsizex", C.c_int),
        ("gpu_binsizey", C.c_int),
        ("gpu_binsizez", C.c_int),
        ("gpu_maxsubprobsize", C.c_int),
        ("gpu_spreadinterponly", C.c_int),
        ("gpu_device_id", C.c_int),
        ("gpu_stream", C.c_void_p),
    ]


class B2nPlanInfo(C.Structure):
    _fields_ = [
        ("type", C.c_int), ("dim", C.c_int), ("is_double", C.c_int), ("ns", C.c_int),
        ("method", C.c_int), ("ntransf", C.c_int), ("batchsize", C.c_int), ("ncoef", C.c_int),
        ("beta", C.c_double), ("upsampfac", C.c_double),
        ("nf", C.c_int64 * 3), ("ms", C.c_int64 * 3),
        ("binsize", C.c_int * 3), ("nbins", C.c_int * 3),
        ("M", C.c_int64), ("N", C.c_int64),
        ("t3_nf_inner", C.c_int64 * 3),
        ("t3_X", C.c_double * 3), ("t3_C", C.c_double * 3), ("t3_S", C.c_double * 3),
        ("t3_D", C.c_double * 3), ("t3_h", C.c_double * 3), ("t3_gam", C.c_double * 3),
    ]


class B2nFfiAttrs(C.Structure):
    """``b2n_ffi_attrs``: the typed attributes of one custom call (lib/jax_finufft_gpu.cc:28-60)."""

    _fields_ = [("eps", C.c_double)] + [(k, C.c_int64) for k in
                                        ("iflag", "n_tot", "n_transf", "n_j", "n_k_1", "n_k_2", "n_k_3", "modeord")] + \
               [("upsampfac", C.c_double)] + [(k, C.c_int64) for k in
                                              ("gpu_method", "gpu_sort", "gpu_kerevalmeth", "gpu_maxbatchsize", "debug")]


EXPORTED = [
    "b2n_default_opts", "b2n_makeplan", "b2n_setpts", "b2n_execute", "b2n_destroy",
    "b2n_plan_info_get", "b2n_plan_sort_get", "b2n_plan_sort_copy", "b2n_run", "b2n_run_host", "b2n_cache_clear",
    "b2n_plan_timings", "b2n_setup_spreader", "b2n_next235beven", "b2n_set_nf_type12",
    "b2n_fseries", "b2n_horner_table", "b2n_default_binsize", "b2n_version", "b2n_launch_count",
    "b2n_ffi_call", "b2n_ffi_arity", "b2n_ffi_targets", "b2n_strerror", "b2n_set_setpts_cache", "b2n_slab_partition",
    "b2n_set_cache_limit", "b2n_cache_bytes", "b2n_slab_fft_xy", "b2n_slab_fft_z",
    "b2n_stack_scaled", "b2n_grad_points",
]


def build(verbose=False):
    """Compile every CUDA source for sm_100a into ``libb200nufft.so`` (nvcc, in-tree)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j", str(os.cpu_count() or 4)]
    res = subprocess.run(cmd, capture_output=not verbose, text=True)
    if res.returncode != 0:
        raise RuntimeError("building libb200nufft.so failed:\n" + (res.stdout or "") + (res.stderr or ""))


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the B200 NUFFT backend is not built. Run "
                "`make -C jax_finufft_b200/csrc -j` (or `python -c 'import __graft_entry__ as g; g.build()'`). "
                "There is no CPU or pure-PyTorch fallback."
            )
        L = C.CDLL(LIB_PATH)
        vp, i64, dbl, ci = C.c_void_p, C.c_int64, C.c_double, C.c_int
        L.b2n_version.restype = C.c_char_p
        L.b2n_launch_count.restype = C.c_ulonglong
        L.b2n_default_opts.argtypes = [C.POINTER(B2nOpts)]
        L.b2n_default_opts.restype = None
        L.b2n_makeplan.argtypes = [ci, ci, C.POINTER(i64), ci, ci, dbl, ci, C.POINTER(vp), C.POINTER(B2nOpts)]
        L.b2n_setpts.argtypes = [vp, i64, vp, vp, vp, i64, vp, vp, vp]
        L.b2n_execute.argtypes = [vp, vp, vp]
        L.b2n_destroy.argtypes = [vp]
        L.b2n_plan_info_get.argtypes = [vp, C.POINTER(B2nPlanInfo)]
        L.b2n_plan_sort_get.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(i64)]
        L.b2n_plan_sort_copy.argtypes = [vp, vp, vp]
        L.b2n_plan_timings.argtypes = [vp, C.POINTER(dbl)]
        L.b2n_run.argtypes = [ci, ci, ci, vp, dbl, ci, i64, ci, i64, C.POINTER(i64), C.POINTER(B2nOpts), vp,
                              C.POINTER(vp), C.POINTER(vp), vp]
        L.b2n_run_host.argtypes = [ci, ci, ci, dbl, ci, i64, ci, i64, C.POINTER(i64), C.POINTER(B2nOpts), vp,
                                   C.POINTER(vp), C.POINTER(vp), vp]
        L.b2n_ffi_call.argtypes = [C.c_char_p, vp, C.POINTER(B2nFfiAttrs), C.POINTER(vp), ci, vp]
        L.b2n_ffi_arity.argtypes = [C.c_char_p]
        L.b2n_ffi_targets.restype = C.POINTER(C.c_char_p)
        L.b2n_strerror.argtypes = [ci]
        L.b2n_strerror.restype = C.c_char_p
        L.b2n_cache_clear.restype = None
        L.b2n_set_cache_limit.argtypes = [C.c_longlong]
        L.b2n_set_cache_limit.restype = C.c_longlong
        L.b2n_cache_bytes.argtypes = [C.POINTER(C.c_ulonglong), C.POINTER(C.c_ulonglong)]
        L.b2n_cache_bytes.restype = None
        L.b2n_set_setpts_cache.argtypes = [ci]
        L.b2n_slab_partition.argtypes = [ci, vp, i64, vp, vp, vp, vp, i64, ci, ci, vp, vp, vp, vp, vp]
        L.b2n_set_setpts_cache.restype = ci
        L.b2n_slab_fft_xy.argtypes = [ci, vp, vp, i64, i64, i64, i64, i64, ci, ci, ci, ci, dbl, vp]
        L.b2n_stack_scaled.argtypes = [ci, vp, i64, ci, i64, ci, vp, C.POINTER(vp), vp]
        L.b2n_grad_points.argtypes = [ci, vp, i64, ci, i64, ci, ci, ci, ci, dbl, vp, vp, vp]
        L.b2n_slab_fft_z.argtypes = [ci, vp, vp, i64, i64, i64, ci, ci, ci, dbl, vp]
        L.b2n_setup_spreader.argtypes = [dbl, dbl, ci, ci, C.POINTER(ci), C.POINTER(dbl)]
        L.b2n_next235beven.argtypes = [i64, i64]
        L.b2n_next235beven.restype = i64
        L.b2n_set_nf_type12.argtypes = [i64, dbl, ci]
        L.b2n_set_nf_type12.restype = i64
        L.b2n_fseries.argtypes = [i64, ci, dbl, C.POINTER(dbl)]
        L.b2n_fseries.restype = None
        L.b2n_horner_table.argtypes = [ci, dbl, ci, C.POINTER(dbl)]
        L.b2n_default_binsize.argtypes = [ci, ci, ci, ci, C.POINTER(ci)]
        L.b2n_default_binsize.restype = None
        _lib = L
    return _lib


def default_opts():
    o = B2nOpts()
    lib().b2n_default_opts(C.byref(o))
    return o
