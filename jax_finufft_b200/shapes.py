"""Operand canonicalisation for the custom call (the job of the reference's
``src/jax_finufft/shapes.py``, on torch tensors).

The backend sees exactly one layout: ``source (n_tot, n_transf, payload...)`` and every point array
``(n_tot, M)``.  A user call may carry any number of leading dimensions on either side, related
by NumPy broadcasting.  The rule that maps one onto the other:

* a leading axis along which the POINTS vary belongs to ``n_tot``: a new point set, hence a new
  bin-sort (setpts) per index;
* a leading axis along which the points are constant (absent or of length 1) while the source
  varies belongs to ``n_transf``: stacked transforms that share one sorted point set -- the
  many-vector fast path (V/include/cufinufft/impl.h:123-127 batches them).

``Folded.unflatten`` undoes the folding on the result.  The attribute names ``broadcast_from`` /
``broadcast_to`` / ``expected_output_shape`` are the reference's (tests/shapes_test.py reads
them), so its test cases run against this module unchanged.
"""

from math import prod

import torch

__all__ = ["abstract_eval", "broadcast_and_flatten_inputs", "Folded"]


class Folded:
    """How a result of shape (n_tot, n_transf, out...) maps back to the caller's leading axes."""

    __slots__ = ("broadcast_from", "broadcast_to", "expected_output_shape")

    def __init__(self, shared_axes, parked_axes, out_shape):
        self.broadcast_from = tuple(shared_axes)      # caller's positions of the shared-points axes
        self.broadcast_to = tuple(parked_axes)        # where they sit while folded (end of the lead block)
        self.expected_output_shape = tuple(out_shape)

    def unflatten(self, result):
        full = result.reshape(self.expected_output_shape)
        if not self.broadcast_to:
            return full
        return torch.movedim(full, self.broadcast_to, self.broadcast_from)


def _common_lead(arrays):
    """Broadcast arrays of shape (lead..., n) against each other; returns (arrays, lead, n)."""
    arrays = torch.broadcast_tensors(*arrays)
    shape = tuple(arrays[0].shape)
    return list(arrays), shape[:-1], shape[-1]


def broadcast_and_flatten_inputs(nufft_type, output_shape, source, *points):
    ndim = len(points) // 2 if nufft_type == 3 else len(points)
    if ndim < 1:
        raise AssertionError("at least one coordinate array is required")
    groups = [list(points[:ndim])] + ([list(points[ndim:])] if nufft_type == 3 else [])

    # 1. one leading shape for all coordinate arrays (type 3: sources and targets together)
    folded_groups, lengths, leads = [], [], []
    for g in groups:
        g, lead, n = _common_lead(g)
        folded_groups.append(g)
        leads.append(lead)
        lengths.append(n)
    lead = tuple(torch.broadcast_shapes(*leads))
    folded_groups = [[p.broadcast_to(lead + (n,)) for p in g] for g, n in zip(folded_groups, lengths)]

    # 2. payload = the trailing axes of `source` that are data, not batch
    payload_rank = ndim if nufft_type == 2 else 1
    # a source with exactly one more batch axis than the points stacks transforms on shared points
    if source.ndim - payload_rank == len(lead) + 1:
        lead = lead + (1,)
        folded_groups = [[p.unsqueeze(-2) for p in g] for g in folded_groups]
    nlead = len(lead)

    # 3. which leading axes do the points share?
    full = tuple(torch.broadcast_shapes(tuple(source.shape[:nlead]), lead))
    shared = [ax for ax in range(nlead) if lead[ax] != full[ax]]
    if any(lead[ax] != 1 for ax in shared):
        raise AssertionError("points can only be shared along axes where they have length 1")
    parked = list(range(nlead - len(shared), nlead))

    source = source.broadcast_to(full + tuple(source.shape[nlead:]))
    if shared:
        source = torch.movedim(source, shared, parked)
        folded_groups = [[torch.movedim(p, shared, parked) for p in g] for g in folded_groups]

    # 4. fold: (own-points axes) -> n_tot, (shared-points axes) -> n_transf
    keep = nlead - len(shared)
    n_tot = prod(source.shape[:keep])
    n_transf = prod(source.shape[keep:nlead])
    moved_lead = tuple(source.shape[:nlead])
    payload = tuple(source.shape[nlead:])
    if len(payload) != payload_rank:
        raise AssertionError(f"source has {len(payload)} data axes, the transform needs {payload_rank}")
    if nufft_type == 2:
        out_tail = (lengths[0],)
    else:
        if payload[0] != lengths[0]:
            raise AssertionError("source and points disagree on the number of non-uniform points")
        if nufft_type == 1:
            if output_shape is None:
                raise AssertionError("type 1 needs output_shape")
            out_tail = tuple(output_shape)
        else:
            out_tail = (lengths[1],)

    source = source.reshape((n_tot, n_transf) + payload)
    flat = [p.reshape(n_tot, n) for g, n in zip(folded_groups, lengths) for p in g]
    return (Folded(shared, parked, moved_lead + out_tail), source, *flat)


def abstract_eval(source, *points, output_shape, nufft_type, **_):
    """(shape, dtype) of the custom call's result for canonical operands, after checking the
    operand contract the C++ side relies on (it receives untyped buffers,
    lib/jax_finufft_gpu.cc:66-192)."""
    if nufft_type not in (1, 2, 3):
        raise ValueError("nufft_type must be 1, 2, or 3")
    ndim = len(points) // 2 if nufft_type == 3 else len(points)
    assert 1 <= ndim <= 3, "1 to 3 dimensions"
    real = {torch.complex64: torch.float32, torch.complex128: torch.float64}.get(source.dtype)
    assert real is not None and all(p.dtype == real for p in points), \
        "source must be complex64/complex128 with matching float32/float64 points"
    src_pts, tgt_pts = points[:ndim], points[ndim:]
    n_tot = source.shape[0]
    for grp in (src_pts, tgt_pts):
        assert all(p.ndim == 2 and p.shape == grp[0].shape for p in grp), "point arrays must be (n_tot, n), all alike"
    assert src_pts[0].shape[0] == n_tot, "source and points disagree on n_tot"
    head = tuple(source.shape[:2])
    if nufft_type == 2:
        assert source.ndim == 2 + ndim, "type 2 source must be (n_tot, n_transf, modes...)"
        return head + (src_pts[0].shape[1],), source.dtype
    assert source.ndim == 3, "source must be (n_tot, n_transf, n_points)"
    if nufft_type == 1:
        assert source.shape[2] == src_pts[0].shape[1], "source and points disagree on the number of points"
        return head + tuple(output_shape), source.dtype
    assert tgt_pts[0].shape[0] == n_tot, "targets and sources disagree on n_tot"
    return head + (tgt_pts[0].shape[1],), source.dtype
