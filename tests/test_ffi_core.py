"""The custom-call boundary below the XLA shim (csrc/ffi_core.cpp: b2n_ffi_call): target names,
arity and argument validation on the CPU; one call per target family on the GPU."""
import ctypes as C

import numpy as np
import pytest

from jax_finufft_b200 import _lib, lowering

# keys of jax_finufft_gpu.registrations() upstream (lib/jax_finufft_gpu.cc:391-420)
UPSTREAM = [f"nufft{d}d{t}{p}" for t in (1, 2, 3) for d in (1, 2, 3) for p in ("f", "")]


def test_targets_are_the_reference_registrations():
    names = lowering.registrations()
    assert sorted(names) == sorted(UPSTREAM) and len(names) == 18
    # the shim relies on single/double alternating in the table
    assert all(n.endswith("f") == (i % 2 == 0) for i, n in enumerate(names))


def test_arity_matches_reference_bindings():
    # 1 + dim operands for types 1/2, 1 + 2*dim for type 3 (lib/jax_finufft_gpu.cc:66-192)
    L = _lib.lib()
    for n in UPSTREAM:
        d, t = int(n[5]), int(n[7])
        assert L.b2n_ffi_arity(n.encode()) == 1 + (2 * d if t == 3 else d), n
    for bad in (b"nufft4d1", b"nufft1d4", b"nufft1d1g", b"nufft", b"", b"fft1d1f", b"nufft1x1"):
        assert L.b2n_ffi_arity(bad) == -1


def test_invalid_calls_are_rejected_before_device_work():
    L = _lib.lib()
    a = _lib.B2nFfiAttrs()
    a.eps, a.iflag, a.n_tot, a.n_transf, a.n_j, a.n_k_1, a.upsampfac = 1e-6, 1, 1, 1, 10, 8, 2.0
    buf = (C.c_double * 64)()
    ptr = C.cast(buf, C.c_void_p)
    two = (C.c_void_p * 2)(ptr, ptr)
    three = (C.c_void_p * 3)(ptr, ptr, ptr)
    assert L.b2n_ffi_call(b"nufft9d1f", None, C.byref(a), two, 2, ptr) == 21      # unknown target
    assert L.b2n_ffi_call(b"nufft1d1f", None, C.byref(a), three, 3, ptr) == 21    # wrong operand count
    assert L.b2n_ffi_call(b"nufft1d1f", None, None, two, 2, ptr) == 21            # no attributes
    assert L.b2n_ffi_call(b"nufft1d1f", None, C.byref(a), two, 2, None) == 21     # no result buffer
    nul = (C.c_void_p * 2)(ptr, None)
    assert L.b2n_ffi_call(b"nufft1d1f", None, C.byref(a), nul, 2, ptr) == 21      # null operand
    a.n_transf = 0
    assert L.b2n_ffi_call(b"nufft1d1f", None, C.byref(a), two, 2, ptr) == 9       # n_transf < 1


def test_strerror_covers_every_code():
    L = _lib.lib()
    for code in (0, 1, 2, 7, 8, 9, 10, 11, 12, 14, 15, 16, 17, 18, 19, 20, 21):
        msg = L.b2n_strerror(code).decode()
        assert msg and "unknown" not in msg
    assert "unknown" in L.b2n_strerror(99).decode()


def test_ffi_attributes_schema_matches_reference():
    # names and order-free typed attributes of lib/jax_finufft_gpu.cc:28-60
    attrs = lowering.ffi_attributes((2, 3, 100), [(2, 100)] * 2, output_shape=(12, 14), iflag=1, eps=1e-6,
                                    opts=None, nufft_type=1, single=True)
    assert set(attrs) == {"eps", "iflag", "n_tot", "n_transf", "n_j", "n_k_1", "n_k_2", "n_k_3", "modeord",
                          "upsampfac", "gpu_method", "gpu_sort", "gpu_kerevalmeth", "gpu_maxbatchsize", "debug"}
    assert set(attrs) == {f[0] for f in _lib.B2nFfiAttrs._fields_}
    # dims reversed: n_k_1 is the LAST JAX axis (lowering.py:96-105)
    assert (attrs["n_k_1"], attrs["n_k_2"], attrs["n_k_3"]) == (14, 12, 0)
    assert (attrs["n_tot"], attrs["n_transf"], attrs["n_j"]) == (2, 3, 100)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ["nufft1d1f", "nufft2d2", "nufft3d3f", "nufft2d1", "nufft3d2f", "nufft1d3"])
def test_ffi_call_against_nudft(name):
    import torch

    L = _lib.lib()
    dim, typ, single = int(name[5]), int(name[7]), name.endswith("f")
    rd, cd = (np.float32, np.complex64) if single else (np.float64, np.complex128)
    eps = 1e-5 if single else 1e-9
    rng = np.random.default_rng(11)
    M, nk = 300, [10, 12, 8][:dim]
    x = [rng.uniform(-np.pi, np.pi, M).astype(rd) for _ in range(dim)]       # x[0] = fastest grid axis
    ks = np.meshgrid(*[np.arange(-(n // 2), (n + 1) // 2) for n in nk[::-1]], indexing="ij")[::-1]
    N3 = 40
    s = [rng.uniform(-5, 5, N3).astype(rd) for _ in range(dim)]
    a = _lib.B2nFfiAttrs()
    a.eps, a.iflag, a.n_tot, a.n_transf, a.n_j, a.upsampfac, a.gpu_sort, a.gpu_kerevalmeth = eps, 1, 1, 1, M, 2.0, 1, 1
    if typ == 3:
        a.n_k_1 = N3
    else:
        a.n_k_1, a.n_k_2, a.n_k_3 = (nk + [0, 0])[:3]
    if typ == 2:
        src = (rng.normal(size=nk[::-1]) + 1j * rng.normal(size=nk[::-1])).astype(cd)
        phase = sum(k[..., None] * xx for k, xx in zip(ks, x))               # (modes..., M)
        expect = np.tensordot(src, np.exp(1j * phase), axes=dim)
    else:
        src = (rng.normal(size=M) + 1j * rng.normal(size=M)).astype(cd)
        if typ == 1:
            phase = sum(k[..., None] * xx for k, xx in zip(ks, x))
            expect = (np.exp(1j * phase) * src).sum(-1)
        else:
            phase = sum(np.outer(ss, xx) for ss, xx in zip(s, x))
            expect = np.exp(1j * phase) @ src
    dev = [torch.as_tensor(v, device="cuda") for v in [src] + x + (s if typ == 3 else [])]
    out = torch.zeros(expect.shape, dtype=torch.complex64 if single else torch.complex128, device="cuda")
    ops = (C.c_void_p * len(dev))(*[t.data_ptr() for t in dev])
    st = torch.cuda.current_stream().cuda_stream
    rc = L.b2n_ffi_call(name.encode(), C.c_void_p(st), C.byref(a), ops, len(dev), C.c_void_p(out.data_ptr()))
    assert rc == 0
    got = out.cpu().numpy()
    err = np.linalg.norm(got - expect) / np.linalg.norm(expect)
    assert err < 10 * eps, err
