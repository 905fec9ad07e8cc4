"""Known-answer and property checks for any module exposing the
``honeybadgermpc.ntl`` API.  They restate the reference's own tests
(/root/reference/tests/test_ntl.py and friends; line numbers cited per check)
so the same checks run against the CPU oracle (``-m "not gpu"``) and against
the CUDA path (``-m gpu``)."""

import random

from conftest import BLS12_381_R as P
from conftest import ROOTS_OF_UNITY


def _corrupt(rng, message, num_errors, num_nones, max_val=131):
    """tests/test_ntl.py:317-328 (seeded)."""
    message = list(message)
    idx = rng.sample(range(len(message)), num_errors + num_nones)
    for i in idx[:num_errors]:
        message[i] = rng.randint(0, max_val)
    for i in idx[num_errors:]:
        message[i] = None
    return message


def check_small_kats(ntl):
    # tests/test_ntl.py:18-28
    assert ntl.lagrange_interpolate([1, 2], [1, 2], P) == [0, 1]
    # :31-41
    assert ntl.vandermonde_batch_interpolate([1, 2], [[1, 2], [3, 5]], P) == [[0, 1], [1, 2]]
    # :44-54
    assert ntl.vandermonde_batch_evaluate([1, 2], [[0, 1], [1, 2]], P) == [[1, 2], [3, 5]]
    # :57-68
    assert ntl.fft([0, 1], 5, 13, 4) == [1, 5, 12, 8]
    # tests/test_reed_solomon.py:21-24 (points 1..4)
    assert ntl.vandermonde_batch_evaluate([1, 2, 3, 4], [[1, 2]], P) == [[3, 5, 7, 9]]
    assert ntl.vandermonde_batch_evaluate([1, 2, 3, 4], [[1, 2], [2, 3]], P) == [
        [3, 5, 7, 9], [5, 8, 11, 14]]
    # :52-53  z=[1,3] -> x=[2,4]
    assert ntl.vandermonde_batch_interpolate([2, 4], [[5, 9]], P) == [[1, 2]]
    assert ntl.vandermonde_batch_interpolate([2, 4], [[5, 9], [8, 14]], P) == [[1, 2], [2, 3]]


def check_fft_properties(ntl, seed=1):
    rng = random.Random(seed)
    # tests/test_ntl.py:71-87 (test_fft_big)
    d, r = 20, 5
    n = 2 ** r
    omega = ROOTS_OF_UNITY[r]
    coeffs = [rng.randrange(P) for _ in range(d)]
    want = [sum(coeffs[j] * pow(pow(omega, i, P), j, P) for j in range(d)) % P for i in range(n)]
    assert ntl.fft(coeffs, omega, P, n) == want
    # :119-136 (test_partial_fft_big)
    assert ntl.partial_fft(coeffs, omega, P, n, 25) == want[:25]
    # :90-116 (test_fft_batch_evaluate_big)
    batch = [[rng.randrange(P) for _ in range(d)] for _ in range(64)]
    got = ntl.fft_batch_evaluate(batch, omega, P, n, 25)
    assert len(got) == 64
    for row, c in zip(got, batch):
        assert row == [sum(c[j] * pow(pow(omega, i, P), j, P) for j in range(d)) % P
                       for i in range(25)]
    # coefficients beyond n are dropped, not wrapped (rsdecode_impl.h:173-175)
    long = [rng.randrange(P) for _ in range(12)]
    w8 = ROOTS_OF_UNITY[3]
    assert ntl.fft(long, w8, P, 8) == ntl.fft(long[:8], w8, P, 8)


def check_fft_interpolate(ntl):
    # tests/test_ntl.py:139-156
    omega = ROOTS_OF_UNITY[3]
    n = 8
    zs = [3, 0]
    xs = [pow(omega, z, P) for z in zs]
    poly = [1, 2]
    ys = [sum(poly[i] * pow(x, i, P) for i in range(2)) % P for x in xs]
    assert ntl.fft_interpolate(zs, ys, omega, P, n) == poly
    # :159-179
    zs = [3, 0, 5]
    xs = [pow(omega, z, P) for z in zs]
    polys = [[1, 2, 0], [3, 2, 1], [3, 4, 2]]
    ys = [[sum(q[i] * pow(x, i, P) for i in range(3)) % P for x in xs] for q in polys]
    assert ntl.fft_batch_interpolate(zs, ys, omega, P, n) == polys


def check_evaluate(ntl, seed=2):
    # tests/test_ntl.py:182-193
    rng = random.Random(seed)
    coeffs = [1, 2, 3, 4]
    for _ in range(20):
        x = rng.randrange(P)
        assert ntl.evaluate(coeffs, x, P) == (1 + 2 * x + 3 * x * x + 4 * x ** 3) % P


def _gao_suite(ntl, int_msg, p, use_omega, seed):
    rng = random.Random(seed)
    k, n = len(int_msg), 22
    t = k - 1
    if use_omega:
        omega, order = ROOTS_OF_UNITY[5], 32
        z = list(range(n))
        x = [pow(omega, zi, p) for zi in z]
        kw = dict(z=z, omega=omega, order=order, use_omega_powers=True)
    else:
        x = list(range(n))
        kw = {}
    encoded = [sum(int_msg[j] * pow(x[i], j, p) for j in range(k)) % p for i in range(n)]
    decoded, loc = ntl.gao_interpolate(x, encoded, k, p, **kw)
    assert decoded == int_msg
    assert loc == [1]  # no-error early exit, rsdecode_impl.h:296-301
    cmax = n - 2 * t - 1
    emax = cmax // 2
    for ne, nn in [(0, cmax), (emax, 0), (emax // 2, cmax // 4)]:
        for _ in range(3):
            corrupted = _corrupt(rng, encoded, ne, nn)
            coeffs, _ = ntl.gao_interpolate(x, corrupted, k, p, **kw)
            assert coeffs == int_msg


def check_gao(ntl):
    # tests/test_ntl.py:196-229, :232-265, :268-314
    _gao_suite(ntl, [2, 3, 2, 8, 7, 5, 9, 5], 53, False, 11)
    _gao_suite(ntl, [0] * 8, 53, False, 12)
    _gao_suite(ntl, [2, 3, 2, 8, 7, 5, 9, 5], P, True, 13)
    # too many errors -> (None, None)   (pyx:439)
    x = list(range(22))
    msg = [2, 3, 2, 8, 7, 5, 9, 5]
    enc = [sum(msg[j] * pow(x[i], j, 53) for j in range(8)) % 53 for i in range(22)]
    bad = list(enc)
    for i in range(8):
        bad[i] = (bad[i] + 1 + i) % 53
    res = ntl.gao_interpolate(x, bad, 8, 53)
    assert res == (None, None) or res[0] != msg


def check_sqrt(ntl, seed=0):
    # tests/test_ntl.py:331-341
    rng = random.Random(seed)
    for _ in range(25):
        v = rng.randrange(P)
        sq = v * v % P
        r = ntl.sqrt_mod(sq, P)
        assert r * r % P == sq


def check_threads(ntl):
    # reed_solomon.py:455-459 relies on AvailableNTLThreads echoing the knob
    ntl.SetNumThreads(3)
    assert ntl.AvailableNTLThreads() == 3
    assert ntl.GetMaxThreads() == 3
    ntl.SetNTLNumThreads(2)
    assert ntl.AvailableNTLThreads() == 2
    ntl.SetNumThreads(1)


def check_errors(ntl):
    import pytest

    with pytest.raises(ntl.InterpolationError):
        ntl.vandermonde_batch_interpolate([1, 1], [[1, 2]], P)  # pyx:168-169
    with pytest.raises(AssertionError):
        ntl.lagrange_interpolate([1, 2], [1], P)  # pyx:83
    with pytest.raises(AssertionError):
        ntl.gao_interpolate([1, 2], [1], 1, P)  # pyx:396
    with pytest.raises(OverflowError):
        ntl.evaluate([1, -2], 3, P)  # int.to_bytes of a negative, pyx:20-22
    with pytest.raises(ValueError):
        ntl.lagrange_interpolate([None, 2], [1, 2], P)  # pyx:37-46
    # decimal strings are accepted where py_obj_to_ZZ is used (pyx:37-46)
    assert ntl.lagrange_interpolate(["1", b"2"], [1, "2"], str(P)) == [0, 1]
    # ragged rows are zero padded (pyx:217,232-233)
    assert ntl.vandermonde_batch_evaluate([1, 2], [[5], [1, 2]], P) == [[5, 5], [3, 5]]
    # outputs of vandermonde_batch_interpolate keep trailing zeros (pyx:185-193)
    assert ntl.vandermonde_batch_interpolate([1, 2], [[7, 7]], P) == [[7, 0]]
    # values >= p are reduced (to_ZZ_p)
    assert ntl.vandermonde_batch_evaluate([1 + P], [[P + 3, 2 * P + 1]], P) == [[4]]
    # vandermonde_inverse returns NTL's textual matrix form (pyx:115-132)
    assert ntl.vandermonde_inverse([1, 2], 13) == "[[2 12]\n[12 1]\n]"


def check_golden(ntl, golden, max_n=None):
    """Fixtures produced by the reference's own pure-Python code
    (tests/golden/make_golden.py)."""
    p = golden["modulus"]
    for case in golden["encode"]:
        n, k = case["n"], case["k"]
        if max_n and n > max_n:
            continue
        pts = next(e for e in golden["eval_points"]
                   if e["n"] == n and e["use_omega_powers"] == case["use_omega_powers"]
                   ) if any(e["n"] == n and e["use_omega_powers"] == case["use_omega_powers"]
                            for e in golden["eval_points"]) else None
        if pts is not None:
            assert ntl.vandermonde_batch_evaluate(pts["points"], case["coeffs"], p) == case["encoded"]
            if case["use_omega_powers"]:
                assert ntl.fft_batch_evaluate(case["coeffs"], pts["omega"], p, pts["order"], n) \
                    == case["encoded"]
                assert ntl.fft(case["coeffs"][0], pts["omega"], p, pts["order"])[:n] \
                    == case["encoded"][0]
    for case in golden["interpolate"]:
        n, k = case["n"], case["k"]
        if max_n and n > max_n:
            continue
        pts = next(e for e in golden["eval_points"]
                   if e["n"] == n and e["use_omega_powers"] == case["use_omega_powers"])
        xs = [pts["points"][z] for z in case["zs"]]
        assert ntl.vandermonde_batch_interpolate(xs, [case["ys"]], p) == [case["coeffs"]]
        got = ntl.lagrange_interpolate(xs, case["ys"], p)
        assert got + [0] * (k - len(got)) == case["coeffs"]
        if case["use_omega_powers"]:
            assert ntl.fft_batch_interpolate(case["zs"], [case["ys"], case["ys"]], pts["omega"], p,
                                             pts["order"]) == [case["coeffs"], case["coeffs"]]
            assert ntl.fft_interpolate(case["zs"], case["ys"], pts["omega"], p, pts["order"]) \
                == case["coeffs"]
    for case in golden["fft"]:
        if max_n and case["n"] > max_n:
            continue
        assert ntl.fft(case["coeffs"], case["omega"], p, case["n"]) == case["evals"]
    for case in golden["fnt_decode"]:
        if max_n and case["n"] > max_n:
            continue
        assert ntl.fft_interpolate(case["zs"], case["ys"], case["omega"], p, case["n"]) \
            == case["coeffs"]
