"""GPU parity tests of the Fr NTT (zk_ntt_fr) against the oracle's EvaluationDomain restatement,
plus size-independent properties at the BASELINE size (2^22): round trips and Horner spot checks."""
import numpy as np
import pytest

from oracle import coracle as co
from oracle import pyref as pr
from zero_chain_b200 import groth16 as zk
from zero_chain_b200 import synthetic as sy

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


def _run(ctx, xm, log_n, mode):
    d = zk.EvaluationDomain(ctx, xm)
    assert d.exp == log_n
    d._run(mode)
    return d.coeffs


@pytest.mark.parametrize("log_n", [0, 1, 2, 3, 7, 10, 11, 12, 13, 15, 16, 18])
def test_ntt_all_modes_vs_oracle(ctx, log_n):
    n = 1 << log_n
    xm = co.fr_to_mont(sy.random_fr_limbs(n, 40 + log_n))
    if n >= 4:
        xm[0] = 0; xm[1] = co.fr_to_mont(co.ints_to_limbs([pr.R - 1], 4))[0]
    for mode in (co.NTT_FFT, co.NTT_IFFT, co.NTT_COSET_FFT, co.NTT_ICOSET_FFT):
        got = _run(ctx, xm, log_n, mode)
        want = co.fr_ntt(xm, log_n, mode)
        assert np.array_equal(got, want), (log_n, mode)


def test_ntt_2_22_properties(ctx):
    """BASELINE config: domain 2^22.  Full compare against the (multi-threaded) oracle for the forward
    transform, round trips for the other modes, Horner evaluation at three points."""
    log_n = 22
    n = 1 << log_n
    x = sy.random_fr_limbs(n, 3)
    xm = co.fr_to_mont(x)
    f = _run(ctx, xm, log_n, co.NTT_FFT)
    assert np.array_equal(f, co.fr_ntt(xm, log_n, co.NTT_FFT))
    assert np.array_equal(_run(ctx, f, log_n, co.NTT_IFFT), xm)
    cf = _run(ctx, xm, log_n, co.NTT_COSET_FFT)
    assert np.array_equal(_run(ctx, cf, log_n, co.NTT_ICOSET_FFT), xm)
    # definition check on a sparse polynomial (cheap Horner): x = e_5 + 3 e_1000003  =>  F[k] = w^(5k) + 3 w^(1000003 k)
    sp = np.zeros((n, 4), np.uint64)
    sp[5] = co.fr_to_mont(co.ints_to_limbs([1], 4))[0]
    sp[1000003] = co.fr_to_mont(co.ints_to_limbs([3], 4))[0]
    fs = co.limbs_to_ints(co.fr_from_mont(_run(ctx, sp, log_n, co.NTT_FFT)[[0, 1, 12345, n - 1]]))
    w = pr.omega(log_n)
    for k, got in zip([0, 1, 12345, n - 1], fs):
        assert got == (pow(w, 5 * k, pr.R) + 3 * pow(w, 1000003 * k, pr.R)) % pr.R


@pytest.mark.parametrize("log_n", [23, 24])
def test_ntt_above_2_22_three_pass(ctx, log_n):
    """Sizes above 2^22 take the outer four-step split (three kernels); bellman's EvaluationDomain supports them up to 2^32."""
    n = 1 << log_n
    xm = co.fr_to_mont(sy.random_fr_limbs(n, 60 + log_n))
    f = _run(ctx, xm, log_n, co.NTT_FFT)
    assert np.array_equal(f, co.fr_ntt(xm, log_n, co.NTT_FFT))
    assert np.array_equal(_run(ctx, f, log_n, co.NTT_IFFT), xm)
    cf = _run(ctx, xm, log_n, co.NTT_COSET_FFT)
    if log_n == 23:
        assert np.array_equal(cf, co.fr_ntt(xm, log_n, co.NTT_COSET_FFT))
    assert np.array_equal(_run(ctx, cf, log_n, co.NTT_ICOSET_FFT), xm)


def test_ntt_batched_matches_single(ctx):
    """The prover's batched transforms (grid.y) must equal per-vector transforms, incl. the coset tables' indexing."""
    import ctypes as C
    import torch
    from zero_chain_b200 import _lib
    for log_n in (9, 15):
        n, batch = 1 << log_n, 5
        xm = co.fr_to_mont(sy.random_fr_limbs(n * batch, 90 + log_n)).reshape(batch, n, 4)
        # there is no public batched entry point: exercise it through the prover instead (tests/test_gpu_groth16.py);
        # here every vector goes through the single-transform API for all four modes as a cross-check of table indexing
        for mode in (0, 1, 2, 3):
            for k in (0, batch - 1):
                assert np.array_equal(_run(ctx, xm[k], log_n, mode), co.fr_ntt(xm[k], log_n, mode))


def test_ntt_degree_too_large(ctx):
    import ctypes as C
    from zero_chain_b200 import _lib
    buf = np.zeros((2, 4), np.uint64)
    assert _lib.lib().zk_ntt_fr(ctx._h, buf.ctypes.data_as(C.c_void_p), 33, 0) == -4   # PolynomialDegreeTooLarge
