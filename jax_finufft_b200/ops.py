"""Public API -- host-side mirror of ``src/jax_finufft/ops.py`` on torch tensors.

``nufft1 / nufft2 / nufft3`` keep the reference's names, argument order, keyword defaults and
error behaviour (ops.py:50-155).  Each is one *primitive* (a ``torch.autograd.Function``) with
the same four rules the reference registers on its JAX primitives (ops.py:356-386):

* impl       -> ``lowering.bind``  (one native ``b2n_run`` call = one XLA custom call upstream)
* JVP        -> ``_NufftPrimitive.jvp``      (ops.py:158-277: ONE extra stacked transform of the
                same type whose ``n_transf`` axis carries the point tangents)
* transpose  -> ``_NufftPrimitive.backward`` (ops.py:280-314: type 1 <-> type 2, type 3 with
                sources and targets swapped), extended to the point cotangents that JAX obtains
                by JVP-then-transpose
* batching   -> ``_NufftPrimitive.vmap``     (ops.py:317-353: un-mapped points => the mapped axis
                is folded into ``n_transf`` and shares one bin-sort)

All of them re-enter the public functions, so a gradient is again "a stacked transform sharing
points" -- the many-vector hot path of the backend.  Complex cotangents follow torch's
convention (grad = dL/dRe + i dL/dIm, i.e. the adjoint A^H), which is the conjugate of JAX's.
"""

from functools import reduce

import numpy as np
import torch

from . import lowering, options, shapes

__all__ = ["nufft1", "nufft2", "nufft3"]


def get_frequency_array(n, modeord):  # ops.py:114-123
    if modeord == 0:
        return np.arange(-(n // 2), (n + 1) // 2)
    elif modeord == 1:
        f = np.empty(n, dtype=np.int64)
        f[: (n + 1) // 2] = np.arange(0, (n + 1) // 2)
        f[(n + 1) // 2:] = np.arange(-(n // 2), 0)
        return f
    else:
        raise ValueError(f"Unsupported modeord: {modeord}")


def _modeord_of(opts, nufft_type):
    o = options.unpack_opts(opts, nufft_type, True)
    return int(o.modeord) if isinstance(o, options.Opts) else 0


def _freq(n, modeord, dim, ndim, like):
    """Frequency vector of mode axis `dim`, shaped to broadcast over the trailing `ndim` axes."""
    shape = [1] * ndim
    shape[dim] = -1
    k = torch.as_tensor(get_frequency_array(int(n), modeord), device=like.device)
    return k.to(like.real.dtype).reshape(shape)


# ---------------------------------------------------------------- fused stacks / reductions (csrc/gradstack.cu)
def _stack_scaled(base, scales):
    """torch.stack([s[:, None, :] * base  (or base for s is None)  for s in scales], dim=2) in ONE pass:
    base (n_tot, n_transf, n) complex, each scale (n_tot, n) real.  The operand the JVP / VJP rules hand
    to their stacked transform (ref ops.py:254,269)."""
    if not base.is_cuda or len(scales) > 4:
        return torch.stack([base if sc is None else sc[:, None, :] * base for sc in scales], dim=2)
    import ctypes as C
    from . import _lib

    base = base.contiguous()
    n_tot, n_transf, n = base.shape
    rdt = base.real.dtype
    sc = [None if t is None else t.to(rdt).contiguous() for t in scales]
    out = torch.empty((n_tot, n_transf, len(sc), n), dtype=base.dtype, device=base.device)
    ptrs = (C.c_void_p * len(sc))(*[None if t is None else t.data_ptr() for t in sc])
    with torch.cuda.device(base.device):
        ier = _lib.lib().b2n_stack_scaled(int(base.dtype == torch.complex128),
                                          C.c_void_p(torch.cuda.current_stream(base.device).cuda_stream),
                                          n_tot, n_transf, n, len(sc), C.c_void_p(base.data_ptr()), ptrs,
                                          C.c_void_p(out.data_ptr()))
    if ier:
        raise RuntimeError(f"b2n_stack_scaled failed with code {ier}")
    return out


def _grad_points(c, h, first, count, sign, real_part=False):
    """[sign * (c.conj() * h[:, :, first + k]).imag.sum(dim=1) for k < count] in ONE pass (real_part: .real):
    c (n_tot, n_transf, n), h (n_tot, n_transf, K, n) -> list of (n_tot, n) real tensors."""
    if not c.is_cuda or count > 4:
        part = (lambda z: z.real) if real_part else (lambda z: z.imag)
        return [sign * part(c.conj() * h[:, :, first + k]).sum(dim=1) for k in range(count)]
    import ctypes as C
    from . import _lib

    c, h = c.contiguous(), h.contiguous()
    n_tot, n_transf, n = c.shape
    out = torch.empty((count, n_tot, n), dtype=c.real.dtype, device=c.device)
    with torch.cuda.device(c.device):
        ier = _lib.lib().b2n_grad_points(int(c.dtype == torch.complex128),
                                         C.c_void_p(torch.cuda.current_stream(c.device).cuda_stream), n_tot, n_transf, n,
                                         h.shape[2], int(first), int(count), 1 if real_part else 0, float(sign),
                                         C.c_void_p(c.data_ptr()), C.c_void_p(h.data_ptr()), C.c_void_p(out.data_ptr()))
    if ier:
        raise RuntimeError(f"b2n_grad_points failed with code {ier}")
    return list(out.unbind(0))


class _NufftPrimitive(torch.autograd.Function):
    """nufft{1,2,3}_p of the reference (ops.py:356-386) on canonical operands."""

    generate_vmap_rule = False

    @staticmethod
    def forward(nufft_type, output_shape, iflag, eps, opts, source, *points):
        return lowering.bind(source, *points, output_shape=output_shape, iflag=iflag, eps=eps, opts=opts,
                             nufft_type=nufft_type)

    @staticmethod
    def setup_context(ctx, inputs, output):
        nufft_type, output_shape, iflag, eps, opts, source, *points = inputs
        ctx.params = (nufft_type, output_shape, iflag, eps, opts)
        ctx.save_for_backward(source, *points)
        ctx.save_for_forward(source, *points)

    # ---------------------------------------------------------------- reverse mode (transpose)
    @staticmethod
    def backward(ctx, g):
        nufft_type, output_shape, iflag, eps, opts = ctx.params
        source, *points = ctx.saved_tensors
        need = ctx.needs_input_grad[5:]
        bopts = options.unpack_opts(opts, {1: 2, 2: 1, 3: 3}[nufft_type], False)
        modeord = _modeord_of(opts, nufft_type)
        grad_source = None
        grad_points = [None] * len(points)
        g = g.contiguous()
        exp = lambda p: p[:, None]  # points shared by every stacked transform

        if nufft_type == 1:
            ndim = len(points)
            stack = [g] if need[0] else []
            dims = [d for d in range(ndim) if need[1 + d]]
            for d in dims:
                stack.append(_freq(output_shape[d], modeord, d, ndim, g) * g)
            if stack:
                h = nufft2(torch.stack(stack, dim=2), *map(exp, points), iflag=-iflag, eps=eps, opts=bopts)
                off = 0
                if need[0]:
                    grad_source, off = h[:, :, 0], 1
                if dims:  # dL/dx = iflag * Im(conj(c) * h_d), summed over transforms
                    for d, gp in zip(dims, _grad_points(source, h, off, len(dims), iflag)):
                        grad_points[d] = gp
        elif nufft_type == 2:
            ndim = len(points)
            if need[0]:
                grad_source = nufft1(tuple(source.shape[-ndim:]), g, *points, iflag=-iflag, eps=eps, opts=bopts)
            dims = [d for d in range(ndim) if need[1 + d]]
            if dims:
                fopts = options.unpack_opts(opts, 2, True)
                args = [(1j * iflag) * _freq(source.shape[2 + d], modeord, d, ndim, source) * source for d in dims]
                h = nufft2(torch.stack(args, dim=2), *map(exp, points), iflag=iflag, eps=eps, opts=fopts)
                for d, gp in zip(dims, _grad_points(g, h, 0, len(dims), 1.0, real_part=True)):
                    grad_points[d] = gp
        else:
            ndim = len(points) // 2
            x, s = points[:ndim], points[ndim:]
            dims = [d for d in range(ndim) if need[1 + d]]
            scales = ([None] if need[0] else []) + [s[d] for d in dims]   # the stack [g, s_x g, s_y g, s_z g]
            if scales:
                h = nufft3(_stack_scaled(g, scales), *map(exp, s), *map(exp, x), iflag=-iflag, eps=eps, opts=bopts)
                off = 0
                if need[0]:
                    grad_source, off = h[:, :, 0], 1
                if dims:
                    for d, gp in zip(dims, _grad_points(source, h, off, len(dims), iflag)):
                        grad_points[d] = gp
            tdims = [d for d in range(ndim) if need[1 + ndim + d]]
            if tdims:
                fopts = options.unpack_opts(opts, 3, True)
                h = nufft3(_stack_scaled(source, [x[d] for d in tdims]), *map(exp, x), *map(exp, s), iflag=iflag,
                           eps=eps, opts=fopts)
                # Re(conj(g) * i*iflag * h) = -iflag * Im(conj(g) * h)
                for d, gp in zip(tdims, _grad_points(g, h, 0, len(tdims), -iflag)):
                    grad_points[ndim + d] = gp
        return (None, None, None, None, None, grad_source, *grad_points)

    # ---------------------------------------------------------------- forward mode (ops.py:158-277)
    @staticmethod
    def jvp(ctx, *tangents):
        nufft_type, output_shape, iflag, eps, opts = ctx.params
        source, *points = ctx.saved_tensors
        dsource, *dpoints = tangents[5:]
        modeord = _modeord_of(opts, nufft_type)
        ndim = len(points) // 2 if nufft_type == 3 else len(points)
        exp = lambda p: p[:, None]
        output_tangents, scales, arguments = [], [], []

        if dsource is not None:
            if nufft_type == 2:
                output_tangents.append(_bind(nufft_type, output_shape, iflag, eps, opts, dsource, *points))
            else:
                scales.append(1.0)
                arguments.append(dsource)
        for dim in range(ndim):
            dx = dpoints[dim]
            if dx is None:
                continue
            if nufft_type == 3:
                factor = (1j * iflag * points[ndim + dim])[:, None, :]
            else:
                n = source.shape[-ndim + dim] if nufft_type == 2 else output_shape[dim]
                factor = 1j * iflag * _freq(n, modeord, dim, ndim, source)
            dx = dx[:, None, :]
            if nufft_type == 2:
                scales.append(dx)
                arguments.append(factor * source)
            else:
                scales.append(factor)
                arguments.append(dx * source)
        if nufft_type == 3:
            scales_s, arguments_s = [], []
            for dim in range(ndim):
                dx = dpoints[ndim + dim]
                if dx is None:
                    continue
                factor = (1j * iflag * points[dim])[:, None, :]
                scales_s.append(dx[:, None, :])
                arguments_s.append(factor * source)
            if scales_s:
                t = nufft3(torch.stack(arguments_s, dim=2), *map(exp, points), iflag=iflag, eps=eps, opts=opts)
                output_tangents += [s * t[:, :, n] for n, s in enumerate(scales_s)]
        if scales:
            argument = torch.stack(arguments, dim=2)
            if nufft_type == 3:
                t = nufft3(argument, *map(exp, points), iflag=iflag, eps=eps, opts=opts)
            elif nufft_type == 2:
                t = nufft2(argument, *map(exp, points), iflag=iflag, eps=eps, opts=opts)
            else:
                t = nufft1(tuple(output_shape), argument, *map(exp, points), iflag=iflag, eps=eps, opts=opts)
            output_tangents += [s * t[:, :, n] for n, s in enumerate(scales)]
        if not output_tangents:
            return None
        return reduce(torch.add, output_tangents)

    # ---------------------------------------------------------------- batching (ops.py:317-353)
    @staticmethod
    def vmap(info, in_dims, nufft_type, output_shape, iflag, eps, opts, source, *points):
        bsource, *bpoints = in_dims[5:]
        kwargs = dict(iflag=iflag, eps=eps, opts=opts)
        if all(bx is None for bx in bpoints):
            assert bsource is not None
            source = torch.movedim(source, bsource, 0)
            mapped_points = tuple(p[None] for p in points)
        else:
            num_repeats = info.batch_size
            if bsource is None:
                source = source[None].expand((num_repeats,) + tuple(source.shape))
            else:
                source = torch.movedim(source, bsource, 0)
            mapped_points = []
            for x, bx in zip(points, bpoints):
                if bx is None:
                    mapped_points.append(x[None].expand((num_repeats,) + tuple(x.shape)))
                else:
                    mapped_points.append(torch.movedim(x, bx, 0))
        if nufft_type == 3:
            return nufft3(source, *mapped_points, **kwargs), 0
        elif nufft_type == 2:
            return nufft2(source, *mapped_points, **kwargs), 0
        return nufft1(tuple(output_shape), source, *mapped_points, **kwargs), 0


def _bind(nufft_type, output_shape, iflag, eps, opts, source, *points):
    return _NufftPrimitive.apply(nufft_type, output_shape, iflag, eps, opts, source, *points)


def _check_opts(opts):
    if opts is not None and not isinstance(opts, (options.Opts, options.NestedOpts)):
        raise TypeError("opts must be an Opts, a NestedOpts or None")


def nufft1(output_shape, source, *points, iflag=1, eps=1e-6, opts=None):
    """Type 1 (nonuniform -> uniform): f[k] = sum_j c_j exp(i*iflag*k.x_j).  Mirrors ops.py:50-82."""
    iflag = int(iflag)
    eps = float(eps)
    _check_opts(opts)
    ndim = len(points)
    if not 1 <= ndim <= 3:
        raise ValueError("Only 1-, 2-, and 3-dimensions are supported")
    output_shape = np.atleast_1d(output_shape).astype(np.int64)
    if output_shape.shape != (ndim,):
        raise ValueError(f"output_shape must have shape: ({ndim},)")
    output_shape = tuple(int(n) for n in output_shape)
    index, source, *points = shapes.broadcast_and_flatten_inputs(1, output_shape, source, *points)
    result = _bind(1, output_shape, iflag, eps, opts, source, *points)
    return index.unflatten(result)


def nufft2(source, *points, iflag=-1, eps=1e-6, opts=None):
    """Type 2 (uniform -> nonuniform): c_j = sum_k f_k exp(i*iflag*k.x_j).  Mirrors ops.py:85-111."""
    iflag = int(iflag)
    eps = float(eps)
    _check_opts(opts)
    ndim = len(points)
    if not 1 <= ndim <= 3:
        raise ValueError("Only 1-, 2-, and 3-dimensions are supported")
    index, source, *points = shapes.broadcast_and_flatten_inputs(2, None, source, *points)
    result = _bind(2, None, iflag, eps, opts, source, *points)
    return index.unflatten(result)


def nufft3(source, *points, iflag=-1, eps=1e-6, opts=None):
    """Type 3 (nonuniform -> nonuniform): f_k = sum_j c_j exp(i*iflag*s_k.x_j).  Mirrors ops.py:126-155."""
    iflag = int(iflag)
    eps = float(eps)
    _check_opts(opts)
    twice_ndim = len(points)
    if twice_ndim % 2 != 0:
        raise ValueError("nufft3 requires an even number of point arrays")
    ndim = twice_ndim // 2
    if not 1 <= ndim <= 3:
        raise ValueError("Only 1-, 2-, and 3-dimensions are supported")
    index, source, *points = shapes.broadcast_and_flatten_inputs(3, None, source, *points)
    result = _bind(3, None, iflag, eps, opts, source, *points)
    return index.unflatten(result)
