"""The CPU oracle against the reference's known answers and golden fixtures
(CPU only; this is what pins the checker before it is trusted)."""

import random

import kats
import pytest
from conftest import BLS12_381_R as P

from oracle import hbmpc_oracle as orc


def test_small_kats():
    kats.check_small_kats(orc)


def test_fft_properties():
    kats.check_fft_properties(orc)


def test_fft_interpolate():
    kats.check_fft_interpolate(orc)


def test_evaluate():
    kats.check_evaluate(orc)


def test_gao():
    kats.check_gao(orc)


def test_sqrt():
    kats.check_sqrt(orc)


def test_threads():
    kats.check_threads(orc)


def test_errors():
    kats.check_errors(orc)


def test_golden(golden):
    kats.check_golden(orc, golden)


def test_eval_point_golden(golden):
    # SURVEY.md section 8c golden omegas + fixtures from the reference EvalPoint
    for e in golden["eval_points"]:
        pt = orc.EvalPoint(P, e["n"], e["use_omega_powers"])
        assert pt.order == e["order"]
        assert pt.omega == e["omega"] and pt.omega2 == e["omega2"]
        assert [pt(i) for i in range(e["n"])] == e["points"]
    pt = orc.EvalPoint(P, 16, True)
    assert pt.omega2 == 0x0461237E58FCCED486FA69D8E4E48506E3317AE6451BB89DE69679532AE1234C
    assert pt.omega == 0x1EDC919EC91F38AC5CCD4631F16EDBA4967A6B6CFB0FACA4807B811A823F728D


def test_wb_golden(golden):
    for case in golden["wb"]:
        p, n, k = case["p"], case["n"], case["k"]
        pt = orc.EvalPoint(p, n, case["use_omega_powers"])
        if case["exception"] is not None:
            name, msg = case["exception"].split(":", 1)
            with pytest.raises(Exception) as ei:
                orc.wb_decode(case["received"], n, k, p, pt)
            assert type(ei.value).__name__ == name and str(ei.value) == msg
        else:
            assert orc.wb_decode(case["received"], n, k, p, pt) == case["decoded"], case["label"]
            z = [i for i, v in enumerate(case["received"]) if v is not None]
            coeffs, errs = orc.wb_robust_decode(
                z, [case["received"][i] for i in z], n, k, p, pt)
            assert coeffs == case["decoded"] and errs == case["error_positions"]


def test_robust_decode_kats():
    # tests/test_reed_solomon.py:76-99,168-183: [3,5,0,9] -> ([1,2],[2]) for Gao and WB
    for use_omega in (False, True):
        pt = orc.EvalPoint(P, 4, use_omega)
        enc = [(2 * pt(i) + 1) % P for i in range(4)]
        enc[2] = 0
        assert orc.gao_robust_decode([0, 1, 2, 3], enc, 4, 2, P, pt) == ([1, 2], [2])
        assert orc.wb_robust_decode([0, 1, 2, 3], enc, 4, 2, P, pt) == ([1, 2], [2])


def test_gao_vs_wb_random():
    rng = random.Random(5)
    n, t = 16, 5
    k = t + 1
    pt = orc.EvalPoint(P, n, False)
    for trial in range(4):
        c = [rng.randrange(P) for _ in range(k)]
        enc = [orc.poly_eval(c, pt(i), P) for i in range(n)]
        bad = rng.sample(range(n), trial + 1)
        for i in bad:
            enc[i] = (enc[i] + 1 + rng.randrange(P - 1)) % P
        z = list(range(n))
        g = orc.gao_robust_decode(z, enc, n, k, P, pt)
        w = orc.wb_robust_decode(z, enc, n, k, P, pt)
        assert g == (c, sorted(bad)) and w == (c, sorted(bad))


def test_interpolate_at_zero(golden):
    for case in golden["interpolate_at_zero"]:
        xs = list(range(1, case["t"] + 2))
        coeffs = orc.vandermonde_batch_interpolate(xs, [case["shares"]], P)[0]
        assert coeffs[0] == case["secret"]


def test_sympy_crosscheck():
    """Independent third implementation for the polynomial helpers."""
    from sympy.polys.domains import ZZ
    from sympy.polys.galoistools import gf_div, gf_mul

    rng = random.Random(9)
    a = [rng.randrange(P) for _ in range(9)]
    b = [rng.randrange(P) for _ in range(4)]
    rev = lambda v: [ZZ(x) for x in reversed(v)]  # noqa: E731
    assert orc.poly_mul(a, b, P) == [int(x) for x in reversed(gf_mul(rev(a), rev(b), P, ZZ))]
    q, r = gf_div(rev(a), rev(b), P, ZZ)
    oq, orr = orc.poly_divrem(a, b, P)
    assert oq == [int(x) for x in reversed(q)] and orr == [int(x) for x in reversed(r)]


def test_config3_n64_golden_matches_oracle():
    """tests/golden/wb_n64_v1.json (n = 64, t = 21) is what the oracle's Welch-Berlekamp and Gao
    produce: re-derive three words (the GPU suite checks the kernels against all 36)"""
    import json
    import os

    from conftest import BLS12_381_R as p

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wb_n64_v1.json")) as fh:
        gold = json.load(fh)
    assert gold.get("cross_checked_with_reference") == 2
    for case in (gold["cases"][0], gold["cases"][7], gold["cases"][-1]):
        pt = orc.EvalPoint(p, gold["n"], case["use_omega_powers"])
        word = [int(v, 16) for v in case["received"]]
        want = ([int(v, 16) for v in case["decoded"]], case["errors"])
        z = list(range(gold["n"]))
        assert orc.gao_robust_decode(z, word, gold["n"], gold["t"] + 1, p, pt) == want
    case = gold["cases"][3]
    pt = orc.EvalPoint(p, gold["n"], case["use_omega_powers"])
    assert orc.wb_robust_decode(list(range(64)), [int(v, 16) for v in case["received"]], 64, 22, p, pt) == \
        ([int(v, 16) for v in case["decoded"]], case["errors"])
