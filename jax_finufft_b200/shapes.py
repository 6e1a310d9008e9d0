"""Shape canonicalisation -- host-side mirror of ``src/jax_finufft/shapes.py`` on torch tensors.

``broadcast_and_flatten_inputs`` (shapes.py:27-126) folds every broadcast dimension into the
backend's ``n_transf`` axis and every genuinely batched dimension into ``n_tot``, so the native
call always sees ``source (n_tot, n_transf, .)`` and ``points (n_tot, M)``;
``abstract_eval`` (shapes.py:129-170) is the shape/dtype contract of the primitive.
"""

from dataclasses import dataclass
from typing import Sequence

import numpy as np
import torch

__all__ = ["abstract_eval", "broadcast_and_flatten_inputs"]


@dataclass
class BroadcastIndex:  # shapes.py:13-24
    broadcast_from: Sequence[int]
    broadcast_to: Sequence[int]
    expected_output_shape: Sequence[int]

    def unflatten(self, result):
        out = result.reshape(tuple(self.expected_output_shape))
        if len(self.broadcast_to):
            out = torch.movedim(out, tuple(self.broadcast_to), tuple(self.broadcast_from))
        return out


def broadcast_and_flatten_inputs(nufft_type, output_shape, source, *points):
    if nufft_type == 3:
        num_dim = len(points) // 2
        points3 = points[num_dim:]
        points = points[:num_dim]
    else:
        num_dim = len(points)
    assert num_dim

    points = torch.broadcast_tensors(*points)
    *input_shape, num_points = points[0].shape

    if nufft_type == 3:
        points3 = torch.broadcast_tensors(*points3)
        *input_shape3, num_points3 = points3[0].shape
        input_shape = list(torch.broadcast_shapes(tuple(input_shape), tuple(input_shape3)))
        points = [p.broadcast_to(tuple(input_shape) + (num_points,)) for p in points]
        points3 = [p.broadcast_to(tuple(input_shape) + (num_points3,)) for p in points3]
    else:
        points3 = []

    # Handle unpadded points (shapes.py:56-62)
    if (nufft_type == 2 and source.ndim == len(input_shape) + num_dim + 1) or (
        nufft_type in (1, 3) and source.ndim == len(input_shape) + 2
    ):
        input_shape = tuple(input_shape) + (1,)
        points = tuple(p[..., None, :] for p in points)
        points3 = tuple(p[..., None, :] for p in points3)

    input_shape = tuple(input_shape)
    target_shape = tuple(torch.broadcast_shapes(tuple(source.shape[: len(input_shape)]), input_shape))

    broadcast_from = tuple(
        n for n, (input_dim, target_dim) in enumerate(zip(input_shape, target_shape)) if input_dim != target_dim
    )
    broadcast_to = tuple(len(target_shape) - len(broadcast_from) + n for n in range(len(broadcast_from)))
    assert all(input_shape[n] == 1 for n in broadcast_from)

    source = source.broadcast_to(target_shape + tuple(source.shape[len(target_shape):]))

    if len(broadcast_to):
        source = torch.movedim(source, broadcast_from, broadcast_to)
        points = tuple(torch.movedim(p, broadcast_from, broadcast_to) for p in points)
        points3 = tuple(torch.movedim(p, broadcast_from, broadcast_to) for p in points3)

    num_in = len(target_shape)
    num_axes = num_in - len(broadcast_from)
    size_in = int(np.prod(source.shape[:num_axes], dtype=int))
    size_bcast = int(np.prod(source.shape[num_axes:num_in], dtype=int))

    if nufft_type == 3:
        assert source.ndim == len(target_shape) + 1
        assert source.shape[-1] == num_points
        expected_output_shape = tuple(source.shape[:num_in]) + (num_points3,)
        source_extra_shape = (num_points,)
    elif nufft_type == 2:
        assert source.ndim == num_in + num_dim
        expected_output_shape = tuple(source.shape[:num_in]) + (num_points,)
        source_extra_shape = tuple(source.shape[num_in:])
    elif nufft_type == 1:
        assert source.ndim == len(target_shape) + 1
        assert source.shape[-1] == num_points
        assert output_shape is not None
        expected_output_shape = tuple(source.shape[:num_in]) + tuple(output_shape)
        source_extra_shape = (num_points,)

    source = source.reshape((size_in, size_bcast) + source_extra_shape)
    points = tuple(p.reshape(size_in, num_points) for p in points)
    points3 = tuple(p.reshape(size_in, num_points3) for p in points3)

    return (
        BroadcastIndex(
            broadcast_from=broadcast_from,
            broadcast_to=broadcast_to,
            expected_output_shape=expected_output_shape,
        ),
        source,
        *points,
        *points3,
    )


def abstract_eval(source, *points, output_shape, nufft_type, **_):
    """Output (shape, dtype) of the primitive; asserts the operand contract (shapes.py:129-170)."""
    ndim = len(points) // 2 if nufft_type == 3 else len(points)
    assert 1 <= ndim <= 3

    single = source.dtype == torch.complex64 and all(x.dtype == torch.float32 for x in points)
    double = source.dtype == torch.complex128 and all(x.dtype == torch.float64 for x in points)
    assert single or double, "source must be complex64/complex128 with matching float32/float64 points"

    assert all(p.ndim == 2 for p in points)
    assert all(p.shape == points[0].shape for p in points[1:ndim])
    assert source.shape[0] == points[0].shape[0]

    if nufft_type == 3:
        assert source.ndim == 3
        assert all(p.shape == points[ndim].shape for p in points[ndim + 1:])
        assert all(p.shape[:-1] == p3.shape[:-1] for (p, p3) in zip(points[:ndim], points[ndim:]))
        return tuple(source.shape[:2]) + (points[ndim].shape[-1],), source.dtype
    elif nufft_type == 2:
        assert source.ndim == 2 + ndim
        return tuple(source.shape[:2]) + (points[0].shape[-1],), source.dtype
    elif nufft_type == 1:
        assert source.ndim == 3
        assert source.shape[2] == points[0].shape[1]
        return tuple(source.shape[:2]) + tuple(output_shape), source.dtype
    raise ValueError("nufft_type must be 1, 2, or 3")
