"""Plan arithmetic of libb200nufft (host side, no GPU) against the oracle's restatement of
V/src/cuda/spreadinterp.cpp:48-90, V/src/cuda/common.cu:166-209, V/src/common/utils.cpp."""
import ctypes as C

import numpy as np
import pytest

import oracle
from jax_finufft_b200 import _lib


def _spreader(eps, sigma, is_double, kem=1):
    ns, beta = C.c_int(), C.c_double()
    ier = _lib.lib().b2n_setup_spreader(eps, sigma, kem, int(is_double), C.byref(ns), C.byref(beta))
    return ns.value, beta.value, ier


# values probed from the reference built with nvcc in this image
@pytest.mark.parametrize("eps,sigma,dbl,ns", [
    (1e-4, 2.0, False, 5), (1e-5, 2.0, False, 6), (1e-6, 2.0, False, 7), (1e-7, 2.0, False, 8),
    (1e-4, 1.25, False, 7), (1e-5, 1.25, False, 9), (1e-6, 1.25, False, 10), (1e-7, 1.25, False, 12),
    (1e-6, 2.0, True, 7), (1e-10, 2.0, True, 11), (1e-12, 2.0, True, 13),
])
def test_ns_probed_values(eps, sigma, dbl, ns):
    got, beta, ier = _spreader(eps, sigma, dbl)
    assert got == ns and ier == (1 if (not dbl and eps < 1.2e-7) else 0)  # float eps below machine eps warns
    o_ns, o_beta, o_ier = oracle.setup_spreader(eps, sigma, 1, is_float=not dbl)
    assert (got, ier) == (o_ns, o_ier)
    assert beta == pytest.approx(o_beta, rel=1e-12)


def test_spreader_sweep_matches_oracle():
    for dbl in (False, True):
        for sigma in (2.0, 1.25):
            for eps in 10.0 ** -np.arange(1, 15, 0.5):
                a = _spreader(float(eps), sigma, dbl)
                b = oracle.setup_spreader(float(eps), sigma, 1, is_float=not dbl)
                assert a[0] == b[0] and a[2] == b[2], (eps, sigma, dbl, a, b)
                assert a[1] == pytest.approx(b[1], rel=1e-12)


def test_eps_too_small_is_a_warning():  # spreadinterp.cpp:50-57 -> ier 1, lib/kernels.cc.cu:52 tolerates it
    assert _spreader(1e-9, 2.0, False)[2] == 1
    assert _spreader(1e-17, 2.0, True)[2] == 1


def test_next235beven_and_nf():
    L = _lib.lib()
    for n in list(range(1, 700)) + [1000, 1023, 1025, 4097, 99999, 1 << 20, 3 ** 12 + 1]:
        for b in (1, 2, 4):
            v = L.b2n_next235beven(n, b)
            assert v == oracle.next235beven(n, b)
            assert v >= n and v % 2 == 0 and (n <= 2 or v % b == 0)  # utils.cpp:126: n<=2 returns 2 first
            m = v
            for p in (2, 3, 5):
                while m % p == 0:
                    m //= p
            assert m == 1
    for ms in (1, 2, 7, 50, 256, 1000, 2048, 1 << 20):
        for sigma, ns in ((2.0, 7), (1.25, 10), (2.0, 16), (2.0, 2)):
            assert L.b2n_set_nf_type12(ms, sigma, ns) == oracle.set_nf_type12(ms, sigma, ns)
    assert L.b2n_set_nf_type12(256, 2.0, 7) == 512 and L.b2n_set_nf_type12(2048, 2.0, 6) == 4096


@pytest.mark.parametrize("ns,nf", [(7, 512), (6, 4096), (2, 16), (16, 270), (11, 100)])
def test_kernel_fourier_series(ns, nf):
    beta = oracle.setup_spreader(10.0 ** (1 - ns), 2.0, 1)[1]
    out = np.zeros(nf // 2 + 1)
    _lib.lib().b2n_fseries(nf, ns, beta, out.ctypes.data_as(C.POINTER(C.c_double)))
    ref = oracle.fseries(nf, ns, beta)
    assert np.allclose(out, ref, rtol=1e-12, atol=1e-15 * abs(ref).max())


@pytest.mark.parametrize("dbl", [False, True])
def test_horner_table_fits_es_kernel(dbl):
    """Our piecewise-polynomial table reproduces exp(beta(sqrt(1-z^2)-1)) (spreadinterp.h:64-82) to
    well inside the requested accuracy for every width the backend can select at sigma=2."""
    L = _lib.lib()
    coef = (C.c_double * (24 * 16))()
    for ns in range(2, 17):
        eps = 10.0 ** (1 - ns)
        beta = oracle.setup_spreader(eps, 2.0, 1, is_float=False)[1]
        nc = L.b2n_horner_table(ns, beta, int(dbl), coef)
        assert 3 <= nc <= 24
        tab = np.array(coef[:]).reshape(24, 16)
        worst = 0.0
        for z in np.linspace(-1, 1, 41):
            x1 = (z - ns + 1) / 2.0  # z = 2 x1 + ns - 1
            for j in range(ns):
                v = 0.0
                for k in range(nc):
                    v = v * z + tab[k, j]
                worst = max(worst, abs(v - oracle.es_kernel(x1 + j, ns, beta)))
        floor = 1e-13 if dbl else 2e-8
        assert worst < max(0.3 * eps, floor), (ns, nc, worst)
        assert nc <= ns + 4, (ns, nc)
