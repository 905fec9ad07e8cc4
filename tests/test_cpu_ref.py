"""The C++ CPU restatement (oracle/cpu_ref.cpp: checker #2 and the timed CPU
baseline) against the Python oracle and the reference's known answers."""

import random
import sys

import numpy as np
import pytest
from conftest import BLS12_381_R as P
from conftest import ROOT, ROOTS_OF_UNITY

sys.path.insert(0, ROOT)
import __graft_entry__ as graft  # noqa: E402

from oracle import hbmpc_oracle as orc  # noqa: E402
from oracle import cpu_ref  # noqa: E402


@pytest.fixture(scope="module")
def ref():
    graft.build_oracle()
    return cpu_ref.CpuRef()


def test_kats(ref):
    assert ref.vandermonde_batch_evaluate([1, 2, 3, 4], [[1, 2], [2, 3]], P) == [
        [3, 5, 7, 9], [5, 8, 11, 14]]
    assert ref.vandermonde_batch_interpolate([2, 4], [[5, 9], [8, 14]], P) == [[1, 2], [2, 3]]
    assert ref.fft_batch_evaluate([[0, 1]], 5, 13, 4, 4) == [[1, 5, 12, 8]]
    with pytest.raises(ZeroDivisionError):
        ref.vandermonde_batch_interpolate([1, 1], [[1, 2]], P)


@pytest.mark.parametrize("p", [P, 53, 2 ** 127 - 1])
def test_vandermonde(ref, p):
    rng = random.Random(1)
    for n, d, batch in [(4, 2, 9), (16, 6, 20), (16, 16, 5), (40, 22, 3)]:
        xs = list(range(1, n + 1))
        polys = [[rng.randrange(p) for _ in range(d)] for _ in range(batch)]
        assert ref.vandermonde_batch_evaluate(xs, polys, p) == \
            orc.vandermonde_batch_evaluate(xs, polys, p)
        xk = rng.sample(xs, d)
        assert ref.vandermonde_batch_interpolate(xk, polys, p) == \
            orc.vandermonde_batch_interpolate(xk, polys, p)


@pytest.mark.parametrize("r,d,k,batch", [(1, 2, 2, 2), (3, 5, 8, 4), (4, 6, 16, 33), (5, 20, 25, 8),
                                         (7, 43, 128, 5), (9, 300, 77, 2)])
def test_fft(ref, r, d, k, batch):
    rng = random.Random(r)
    n = 2 ** r
    omega = ROOTS_OF_UNITY[r] if r < len(ROOTS_OF_UNITY) else pow(7, (P - 1) // n, P)
    polys = [[rng.randrange(P) for _ in range(d)] for _ in range(batch)]
    assert ref.fft_batch_evaluate(polys, omega, P, n, k) == orc.fft_batch_evaluate(polys, omega, P, n, k)


@pytest.mark.parametrize("r,k,batch", [(3, 3, 4), (4, 6, 40), (4, 16, 3), (6, 22, 6), (7, 43, 4)])
def test_fft_interpolate(ref, r, k, batch):
    rng = random.Random(r + k)
    n = 2 ** r
    omega = ROOTS_OF_UNITY[r]
    zs = rng.sample(range(n), k)
    ys = [[rng.randrange(P) for _ in range(k)] for _ in range(batch)]
    assert ref.fft_batch_interpolate(zs, ys, omega, P, n) == \
        orc.fft_batch_interpolate(zs, ys, omega, P, n)


def test_golden(ref, golden):
    p = golden["modulus"]
    for case in golden["fft"]:
        assert ref.fft_batch_evaluate([case["coeffs"]], case["omega"], p, case["n"], case["n"])[0] \
            == case["evals"]
    for case in golden["fnt_decode"]:
        assert ref.fft_batch_interpolate(case["zs"], [case["ys"]], case["omega"], p, case["n"])[0] \
            == case["coeffs"]


def test_threads_agree(ref):
    rng = np.random.default_rng(5)
    c = rng.integers(0, 2 ** 63, size=(512, 6, 4), dtype=np.uint64)
    c[:, :, 3] >>= np.uint64(2)
    omega = ROOTS_OF_UNITY[4]
    a = ref.fft_batch_evaluate_limbs(c, omega, P, 16, 16, threads=1)
    b = ref.fft_batch_evaluate_limbs(c, omega, P, 16, 16, threads=0)
    assert np.array_equal(a, b)
