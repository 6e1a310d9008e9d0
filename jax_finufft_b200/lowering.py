"""The primitive's device call -- host-side mirror of ``src/jax_finufft/lowering.py``.

Upstream, ``lowering()`` (lowering.py:28-178) picks the custom-call target
``nufft{dim}d{type}{f}``, packs the FFI attributes, reverses the dimension order (C-order arrays
-> the backend's x-fastest convention, lowering.py:96-105) and emits the XLA custom call that
lands in ``run_nufft`` (lib/kernels.cc.cu:25-92).  Here the same attributes are handed straight
to ``b2n_ffi_call`` (include/b200nufft.h) -- the very function the XLA-FFI shim
(csrc/xla_ffi_shim.cc) calls for each custom call -- which lands in ``b2n_run``.  Operands must be CUDA tensors: there is no CPU path.
"""

import ctypes as C

import numpy as np
import torch

from . import _lib, options

__all__ = ["bind", "op_name", "ffi_attributes", "registrations"]


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


def registrations():
    """Target names the extension module exports (``jax_finufft_gpu.registrations()`` keys,
    lib/jax_finufft_gpu.cc:391-420), read from the library."""
    tg = _lib.lib().b2n_ffi_targets()
    names, i = [], 0
    while tg[i]:
        names.append(tg[i].decode())
        i += 1
    return names


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
    a = _lib.B2nFfiAttrs()
    for k, v in attrs.items():
        setattr(a, k, v)
    n_ops = 1 + (2 if nufft_type == 3 else 1) * ndim
    assert len(operands) == n_ops == L.b2n_ffi_arity(name.encode())
    ops = (C.c_void_p * n_ops)(*[t.data_ptr() for t in operands])
    with torch.cuda.device(out.device):
        stream = torch.cuda.current_stream(out.device).cuda_stream
        ret = L.b2n_ffi_call(name.encode(), C.c_void_p(stream), C.byref(a), ops, n_ops, C.c_void_p(out.data_ptr()))
    if ret > 1:  # 1 = "eps too small" warning, tolerated (lib/kernels.cc.cu:52)
        raise RuntimeError(f"{name}: {L.b2n_strerror(ret).decode()} (code {ret})")
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
