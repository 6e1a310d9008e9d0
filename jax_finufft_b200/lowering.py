"""The primitive's device call -- host-side mirror of ``src/jax_finufft/lowering.py``.

Upstream, ``lowering()`` (lowering.py:28-178) picks the custom-call target
``nufft{dim}d{type}{f}``, packs the FFI attributes, reverses the dimension order (C-order arrays
-> the backend's x-fastest convention, lowering.py:96-105) and emits the XLA custom call that
lands in ``run_nufft`` (lib/kernels.cc.cu:25-92).  Here the same attributes are handed straight
to ``b2n_run`` (include/b200nufft.h), the entry point the XLA-FFI shim
(csrc/xla_ffi_shim.cc) also calls.  Operands must be CUDA tensors: there is no CPU path.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib, options

__all__ = ["bind", "op_name", "ffi_attributes"]


def op_name(ndim, nufft_type, single):
    """Custom-call target name (lowering.py:87-94)."""
    return f"nufft{ndim}d{nufft_type}{'f' if single else ''}"


def ffi_attributes(source_shape, points_shapes, *, output_shape, iflag, eps, opts, nufft_type, single):
    """The attribute dictionary of the custom call (lowering.py:96-174, GPU branch)."""
    ndim = len(points_shapes) // 2 if nufft_type == 3 else len(points_shapes)
    assert 1 <= ndim <= 3
    n_tot, n_transf = source_shape[0], source_shape[1]
    n_j = points_shapes[0][1]
    n_k_full = np.zeros(3, dtype=np.int64)
    if nufft_type == 1:
        n_k_full[:ndim] = np.array(output_shape, dtype=np.int64)[::-1]
    elif nufft_type == 2:
        n_k_full[:ndim] = np.array(source_shape[2:], dtype=np.int64)[::-1]
    else:
        n_k_full[0] = points_shapes[ndim][1]
    if opts is None:
        opts = options.Opts()
    opts = options.unpack_opts(opts, nufft_type, True)
    if opts is None:
        opts = options.Opts()
    assert isinstance(opts, options.Opts)
    native = opts.to_cufinufft_opts()
    return {
        "eps": float(np.float32(eps)) if single else float(eps),
        "iflag": int(iflag),
        "n_tot": int(n_tot),
        "n_transf": int(n_transf),
        "n_j": int(n_j),
        "n_k_1": int(n_k_full[0]),
        "n_k_2": int(n_k_full[1]),
        "n_k_3": int(n_k_full[2]),
        "modeord": native.modeord,
        "upsampfac": native.upsampfac,
        "gpu_method": native.gpu_method,
        "gpu_sort": native.gpu_sort,
        "gpu_kerevalmeth": native.gpu_kerevalmeth,
        "gpu_maxbatchsize": native.gpu_maxbatchsize,
        "debug": native.debug,
    }


def _execute(name, attrs, operands, out):
    """Enqueue one custom call on the current CUDA stream (what the XLA runtime does upstream)."""
    ndim, nufft_type = int(name[5]), int(name[7])
    single = name.endswith("f")
    for t in operands + [out]:
        if not t.is_cuda:
            raise ValueError(
                "jax_finufft_b200 runs on CUDA devices only (sm_100a); got a tensor on "
                f"{t.device}. There is no CPU fallback."
            )
    L = _lib.lib()
    o = _lib.default_opts()
    o.modeord = attrs["modeord"]
    o.upsampfac = attrs["upsampfac"]
    o.gpu_method = attrs["gpu_method"]
    o.gpu_sort = attrs["gpu_sort"]
    o.gpu_kerevalmeth = attrs["gpu_kerevalmeth"]
    o.gpu_maxbatchsize = attrs["gpu_maxbatchsize"]
    o.debug = attrs["debug"]
    n_k = (C.c_int64 * 3)(attrs["n_k_1"], attrs["n_k_2"], attrs["n_k_3"])
    pts = (C.c_void_p * 3)(*[operands[1 + d].data_ptr() for d in range(ndim)] + [None] * (3 - ndim))
    if nufft_type == 3:
        tgt = (C.c_void_p * 3)(*[operands[1 + ndim + d].data_ptr() for d in range(ndim)] + [None] * (3 - ndim))
    else:
        tgt = (C.c_void_p * 3)(None, None, None)
    with torch.cuda.device(out.device):
        stream = torch.cuda.current_stream(out.device).cuda_stream
        ret = L.b2n_run(nufft_type, ndim, 0 if single else 1, C.c_void_p(stream), attrs["eps"], attrs["iflag"],
                        attrs["n_tot"], attrs["n_transf"], attrs["n_j"], n_k, C.byref(o),
                        C.c_void_p(operands[0].data_ptr()), pts, tgt, C.c_void_p(out.data_ptr()))
    if ret > 1:  # 1 = "eps too small" warning, tolerated (lib/kernels.cc.cu:52)
        raise RuntimeError(f"b200nufft {name} failed with code {ret}")
    return out


def bind(source, *points, output_shape, iflag, eps, opts, nufft_type):
    """Evaluate the primitive on canonical operands: source (n_tot, n_transf, ...), points (n_tot, M)."""
    from .shapes import abstract_eval

    ndim = len(points) // 2 if nufft_type == 3 else len(points)
    out_shape, out_dtype = abstract_eval(source, *points, output_shape=output_shape, nufft_type=nufft_type)
    single = source.dtype == torch.complex64
    attrs = ffi_attributes(tuple(source.shape), [tuple(p.shape) for p in points], output_shape=output_shape,
                           iflag=iflag, eps=eps, opts=opts, nufft_type=nufft_type, single=single)
    # Reverse points because the backend uses Fortran order (lowering.py:104-105)
    points_fortran = list(points[:ndim][::-1]) + list(points[ndim:][::-1])
    operands = [source.contiguous()] + [p.contiguous() for p in points_fortran]
    out = torch.empty(out_shape, dtype=out_dtype, device=source.device)
    if out.numel() == 0:
        return out
    return _execute(op_name(ndim, nufft_type, single), attrs, operands, out)
