"""Host-logic tests mirroring the reference's tests/shapes_test.py and tests/options_test.py
(same cases, same expected values) on the torch mirror of shapes.py / options.py."""
import pytest
import torch

from jax_finufft_b200 import options, shapes

E = lambda *s: torch.empty(s)


def test_broadcast_inputs():  # shapes_test.py:6-16
    index, source, *points = shapes.broadcast_and_flatten_inputs(1, (10,), E(50, 5, 1, 7, 12, 5), E(50, 1, 6, 1, 12, 5))
    assert tuple(source.shape) == (50 * 6 * 12, 5 * 7, 5)
    assert len(points) == 1 and tuple(points[0].shape) == (50 * 6 * 12, 5)
    assert index.broadcast_from == (1, 3) and index.broadcast_to == (3, 4)
    assert index.expected_output_shape == (50, 6, 12, 5, 7, 10)


def test_broadcast_inputs_type2():  # shapes_test.py:19-29
    index, source, *points = shapes.broadcast_and_flatten_inputs(2, None, E(50, 5, 1, 7, 12, 11), E(50, 1, 6, 1, 12, 5))
    assert tuple(source.shape) == (50 * 6 * 12, 5 * 7, 11)
    assert tuple(points[0].shape) == (50 * 6 * 12, 5)
    assert index.broadcast_from == (1, 3) and index.broadcast_to == (3, 4)
    assert index.expected_output_shape == (50, 6, 12, 5, 7, 5)


def test_vector_inputs():  # shapes_test.py:32-43
    index, source, *points = shapes.broadcast_and_flatten_inputs(2, None, E(5, 6), E(10), E(10))
    assert tuple(source.shape) == (1, 1, 5, 6)
    assert [tuple(p.shape) for p in points] == [(1, 10), (1, 10)]
    assert index.expected_output_shape == (10,)
    assert index.broadcast_from == () and index.broadcast_to == ()


def test_unpadded_points_type2():  # shapes_test.py:46-57
    index, source, *points = shapes.broadcast_and_flatten_inputs(2, None, E(3, 8, 5, 6), E(3, 10), E(3, 10))
    assert tuple(source.shape) == (3, 8, 5, 6)
    assert [tuple(p.shape) for p in points] == [(3, 10), (3, 10)]
    assert index.expected_output_shape == (3, 8, 10)
    assert index.broadcast_from == (1,) and index.broadcast_to == (1,)


def test_unpadded_points_type1():  # shapes_test.py:60-71
    index, source, *points = shapes.broadcast_and_flatten_inputs(1, (17, 13), E(3, 5, 6), E(1, 6), E(1, 6))
    assert tuple(source.shape) == (1, 15, 6)
    assert [tuple(p.shape) for p in points] == [(1, 6), (1, 6)]
    assert index.expected_output_shape == (3, 5, 17, 13)
    assert index.broadcast_from == (0, 1) and index.broadcast_to == (0, 1)


def test_unpadded_points_type3():  # shapes_test.py:74-
    index, source, *points = shapes.broadcast_and_flatten_inputs(3, None, E(3, 5, 6), E(3, 6), E(3, 6), E(3, 11), E(3, 11))
    assert tuple(source.shape) == (3, 5, 6)
    assert len(points) == 4
    assert [tuple(p.shape) for p in points] == [(3, 6), (3, 6), (3, 11), (3, 11)]
    assert index.expected_output_shape == (3, 5, 11)


def test_abstract_eval_shapes_and_dtype_check():  # shapes.py:129-170
    src = torch.empty((2, 3, 7), dtype=torch.complex64)
    x = torch.empty((2, 7), dtype=torch.float32)
    shp, dt = shapes.abstract_eval(src, x, x, output_shape=(4, 5), nufft_type=1)
    assert tuple(shp) == (2, 3, 4, 5) and dt == torch.complex64
    f = torch.empty((2, 3, 4, 5), dtype=torch.complex128)
    xd = torch.empty((2, 7), dtype=torch.float64)
    shp, dt = shapes.abstract_eval(f, xd, xd, output_shape=None, nufft_type=2)
    assert tuple(shp) == (2, 3, 7) and dt == torch.complex128
    with pytest.raises((ValueError, TypeError, AssertionError)):
        shapes.abstract_eval(src, xd, output_shape=(4,), nufft_type=1)  # c64 with f64 points


@pytest.mark.parametrize("opts", [None, options.Opts()])
def test_default_options(opts):  # options_test.py:5-10
    for t in (1, 2):
        for fwd in (True, False):
            assert options.unpack_opts(opts, t, fwd) == opts


def test_nested_by_type():  # options_test.py:13-20
    opts = options.NestedOpts(type1=options.Opts(spread_debug=True), type2=options.Opts(debug=True))
    assert options.unpack_opts(opts, 1, True) == options.Opts(spread_debug=True)
    assert options.unpack_opts(opts, 2, True) == options.Opts(debug=True)


def test_nested_by_direction():  # options_test.py:23-30
    opts = options.NestedOpts(forward=options.Opts(spread_debug=True), backward=options.Opts(debug=True))
    assert options.unpack_opts(opts, 1, True) == options.Opts(spread_debug=True)
    assert options.unpack_opts(opts, 1, False) == options.Opts(debug=True)


def test_nested_multi():  # options_test.py:33-52
    inner = options.NestedOpts(type1=options.Opts(spread_debug=True, debug=True), type2=options.Opts(debug=True))
    opts = options.NestedOpts(forward=options.Opts(spread_debug=True), backward=inner)
    assert options.unpack_opts(opts, 1, True) == options.Opts(spread_debug=True)
    assert options.unpack_opts(opts, 1, False) == inner
    assert options.unpack_opts(options.unpack_opts(opts, 1, False), 1, True) == options.Opts(spread_debug=True, debug=True)
    assert options.unpack_opts(options.unpack_opts(opts, 1, False), 2, True) == options.Opts(debug=True)


def test_gpu_opts_forwarded_fields():  # options.py:105-119 -- exactly seven fields cross the boundary
    o = options.Opts(modeord=1, gpu_upsampfac=2.0, gpu_method=2, gpu_sort=False, gpu_kerevalmeth=0,
                     gpu_maxbatchsize=4, gpu_debug=True)
    n = o.to_cufinufft_opts()
    assert (n.modeord, n.upsampfac, n.gpu_method, n.gpu_sort, n.gpu_kerevalmeth, n.gpu_maxbatchsize, n.debug) == \
        (1, 2.0, 2, 0, 0, 4, 1)
