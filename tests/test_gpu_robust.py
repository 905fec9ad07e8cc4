"""Parity of the CUDA Gao and Welch-Berlekamp decoders with the oracle and the
reference's known answers (error positions, failure modes, p = 53 corner
cases).  ``pytest -m gpu``."""

import random

import kats
import numpy as np
import pytest
from conftest import BLS12_381_R as P

from oracle import hbmpc_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ntl():
    from honeybadgermpc_b200 import ntl as m

    m._ctx(P)
    return m


@pytest.fixture(scope="module")
def rs():
    from honeybadgermpc_b200 import reed_solomon

    return reed_solomon


def _point(p, n, omega):
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.polynomial import EvalPoint

    return EvalPoint(GF(p), n, omega)


def test_gao_kats(ntl):
    kats.check_gao(ntl)


def _enc(coeffs, xs, p):
    return [orc.poly_eval(coeffs, x, p) for x in xs]


@pytest.mark.parametrize("p", [P, 53, 257])
def test_gao_vs_oracle_random(ntl, p):
    """random words incl. too many errors, erasures and degenerate inputs"""
    rng = random.Random(p % 97)
    for trial in range(60):
        n = rng.choice([4, 7, 10, 16, 22]) if p > 30 else 10
        n = min(n, p - 1)
        k = rng.randint(1, max(1, n // 2))
        xs = rng.sample(range(p), n) if p < 1000 else [rng.randrange(p) for _ in range(n)]
        msg = [rng.randrange(p) for _ in range(k)]
        if trial % 7 == 0:
            msg = [0] * k
        word = _enc(msg, xs, p)
        nerr = rng.randint(0, n - k + 1)
        for i in rng.sample(range(n), min(nerr, n)):
            word[i] = rng.randrange(p)
        for i in rng.sample(range(n), rng.randint(0, max(0, (n - k) // 2))):
            word[i] = None
        if sum(v is not None for v in word) == 0:
            continue
        want = orc.gao_interpolate(xs, word, k, p)
        got = ntl.gao_interpolate(xs, word, k, p)
        assert got == want, (trial, n, k, xs, word)


def test_gao_batch_vs_oracle(ntl):
    from honeybadgermpc_b200 import robust

    rng = random.Random(3)
    for n, k, batch in [(16, 6, 40), (64, 22, 24), (22, 8, 17)]:
        pt = orc.EvalPoint(P, n, True)
        xs = [pt(i) for i in range(n)]
        rows, wants = [], []
        for b in range(batch):
            msg = [rng.randrange(P) for _ in range(k)]
            word = _enc(msg, xs, P)
            nerr = rng.choice([0, 0, 1, (n - k) // 2, (n - k) // 2, (n - k) // 2 + 1, n - k])
            for i in rng.sample(range(n), nerr):
                word[i] = (word[i] + 1 + rng.randrange(P - 1)) % P
            rows.append(word)
            wants.append(orc.gao_interpolate(xs, word, k, P))
        coeffs, locator, loc_len, status = robust.gao_decode_batch_limbs(
            ntl.pack_vec(xs, P), ntl.pack_rows(rows, n, P), k, P)
        ci, li = ntl.unpack_rows(coeffs), ntl.unpack_rows(locator)
        for b in range(batch):
            if wants[b][0] is None:
                assert status[b] == 1
            else:
                assert status[b] == 0
                assert ci[b] == wants[b][0]
                assert li[b][: loc_len[b]] == wants[b][1]


def test_robust_decode_kats(rs):
    # tests/test_reed_solomon.py:76-99,168-183: [3,5,0,9] -> ([1,2],[2]) for Gao and WB
    for use_omega in (False, True):
        pt = _point(P, 4, use_omega)
        enc = [(2 * pt(i).value + 1) % P for i in range(4)]
        enc[2] = 0
        for algo in (rs.Algorithm.GAO, rs.Algorithm.WELCH_BERLEKAMP):
            dec = rs.RobustDecoderFactory.get(1, pt, algo)
            assert dec.robust_decode([0, 1, 2, 3], enc) == ([1, 2], [2])
            assert dec.robust_decode([3, 1, 0, 2], [enc[3], enc[1], enc[0], enc[2]]) == ([1, 2], [2])


def test_wb_golden(rs, golden):
    from honeybadgermpc_b200 import robust

    for case in golden["wb"]:
        p, n, k = case["p"], case["n"], case["k"]
        pt = _point(p, n, case["use_omega_powers"])
        z = [i for i, v in enumerate(case["received"]) if v is not None]
        row = [case["received"][i] for i in z]
        xs = [pt(i).value for i in z]
        if case["exception"] is not None:
            name, msg = case["exception"].split(":", 1)
            if msg == "found no divisors!":
                assert robust.wb_decode_rows(xs, [row], n, n - len(z), k, p) == [None]
            else:
                # per-row outcome: the exception travels with the row (it is raised by
                # whoever consumes the row, reed_solomon.py:334-348) ...
                (got,) = robust.wb_decode_rows(xs, [row], n, n - len(z), k, p)
                assert isinstance(got, BaseException)
                assert str(got) == msg or type(got).__name__ == name
                # ... and the one-row API raises it, like the reference
                with pytest.raises(Exception) as ei:
                    rs.WelchBerlekampRobustDecoder(k - 1, pt).robust_decode(z, row)
                assert str(ei.value) == msg or type(ei.value).__name__ == name
        else:
            assert robust.wb_decode_rows(xs, [row], n, n - len(z), k, p) == [case["decoded"]], case["label"]
            dec = rs.WelchBerlekampRobustDecoder(k - 1, pt)
            assert dec.robust_decode(z, row) == (case["decoded"], case["error_positions"])


def _wb_outcome(fn):
    try:
        return ("ok", fn())
    except AssertionError:
        raise
    except Exception as e:  # noqa: BLE001
        return ("raise", type(e).__name__, str(e))


@pytest.mark.parametrize("p", [53, 257, P])
def test_wb_vs_oracle_random(rs, p):
    """WB incl. beyond-capacity words and tiny fields, where the reference's
    syntactic pivot test and its failure modes matter"""
    rng = random.Random(p % 89 + 1)
    trials = 150 if p < 1000 else 40
    for trial in range(trials):
        n = rng.choice([4, 5, 7, 10, 13, 16])
        t = rng.randint(0, (n - 1) // 3)
        k = t + 1
        use_omega = p == P and trial % 2 == 0
        pt = _point(p, n, use_omega)
        opt = orc.EvalPoint(p, n, use_omega)
        msg = [rng.randrange(p) for _ in range(k)]
        if trial % 9 == 0:
            msg = [0] * k
        xs_all = [opt(i) for i in range(n)]
        word = _enc(msg, xs_all, p)
        for i in rng.sample(range(n), rng.randint(0, min(n, t + 2))):
            word[i] = rng.randrange(p)
        c_max = n - 2 * t - 1
        erase = rng.sample(range(n), rng.randint(0, c_max))
        z = [i for i in range(n) if i not in erase]
        rng.shuffle(z)
        row = [word[i] for i in z]
        want = _wb_outcome(lambda: orc.wb_robust_decode(z, row, n, k, p, opt))
        dec = rs.WelchBerlekampRobustDecoder(t, pt)
        got = _wb_outcome(lambda: dec.robust_decode(z, row))
        if want[0] == "raise":
            assert got[0] == "raise" and got[2] == want[2], (trial, want, got)
        else:
            assert got == want, (trial, n, k, z, row, want, got)


def test_config3_n64_golden(rs):
    """BASELINE configs[2]: n = 64, t = 21, up to t corrupted evaluations per word -- 36 words
    decoded by the oracle's restatement of the reference's Welch-Berlekamp solver (about a
    second per word; tests/golden/make_wb_n64_golden.py, cross-checked there against the
    reference's own class) vs wb_kernel and gao_kernel, batched."""
    import json
    import os

    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wb_n64_v1.json")
    with open(path) as fh:
        gold = json.load(fh)
    n, t = gold["n"], gold["t"]
    assert (n, t) == (64, 21) and len(gold["cases"]) >= 32
    for omega in (False, True):
        cases = [c for c in gold["cases"] if c["use_omega_powers"] == omega]
        pt = _point(P, n, omega)
        rows = [[int(v, 16) for v in c["received"]] for c in cases]
        want = [([int(v, 16) for v in c["decoded"]], c["errors"]) for c in cases]
        z = list(range(n))
        assert rs.WelchBerlekampRobustDecoder(t, pt).robust_decode_batch(z, rows) == want
        assert rs.GaoRobustDecoder(t, pt).robust_decode_batch(z, rows) == want
        # a permuted arrival order of the parties changes nothing
        perm = list(range(n))
        random.Random(4).shuffle(perm)
        prow = [[r[i] for i in perm] for r in rows[:6]]
        assert rs.WelchBerlekampRobustDecoder(t, pt).robust_decode_batch(perm, prow) == want[:6]
        assert rs.GaoRobustDecoder(t, pt).robust_decode_batch(perm, prow) == want[:6]


def test_wb_shortcut_equals_exact_elimination(rs):
    """hbg_wb_decode_batch: the unique-decoding shortcut (Gao kernel for the words within the
    decoding radius, exact elimination kernel for the rest) returns exactly what the exact
    kernel returns for every word -- decodable, beyond capacity (all three failure modes),
    all-zero -- with and without erasures (the shortcut only applies when 2 e_max + k <= m)."""
    from honeybadgermpc_b200 import robust
    from honeybadgermpc_b200.ntl import pack_rows, pack_vec

    rng = random.Random(99)
    for p in (P, 257):
        ctx = robust._ctx(p)
        for n, t in ((4, 1), (7, 2), (10, 3), (16, 5), (13, 4)):
            k = t + 1
            for erased in (0, 1, 2):
                m = n - erased
                if 2 * t + 1 + erased > n:
                    continue
                e_max = (m - t) // 2
                if e_max < 1:
                    continue
                xs = [i + 1 for i in range(n)][:m]
                words = []
                for w in range(60):
                    msg = [rng.randrange(p) for _ in range(k)] if w % 7 else [0] * k
                    word = [orc.poly_eval(msg, x, p) for x in xs]
                    for i in rng.sample(range(m), rng.randint(0, min(m, e_max + 2))):
                        word[i] = rng.randrange(p)
                    words.append(word)
                xl, yl = pack_vec(xs, p), pack_rows(words, m, p)
                ctx.set_wb_path("exact")
                want = robust.wb_decode_batch_limbs(xl, yl, k, e_max, p)
                assert ctx.last_kernel() == "wb_kernel"
                ctx.set_wb_path("auto")
                got = robust.wb_decode_batch_limbs(xl, yl, k, e_max, p)
                for a, b in zip(got, want):
                    assert np.array_equal(a, b), (p, n, t, erased)
                assert set(want[2].tolist()) <= {0, 1, 2, 3}
        ctx.set_wb_path("auto")


def test_gao_equals_wb_when_decodable(rs):
    rng = random.Random(8)
    n, t = 16, 5
    for use_omega in (False, True):
        pt = _point(P, n, use_omega)
        gao = rs.GaoRobustDecoder(t, pt)
        wb = rs.WelchBerlekampRobustDecoder(t, pt)
        rows, bads = [], []
        for trial in range(12):
            c = [rng.randrange(P) for _ in range(t + 1)]
            enc = [orc.poly_eval(c, pt(i).value, P) for i in range(n)]
            bad = sorted(rng.sample(range(n), trial % (t + 1)))
            for i in bad:
                enc[i] = (enc[i] + 1 + rng.randrange(P - 1)) % P
            rows.append(enc)
            bads.append((c, bad))
        z = list(range(n))
        assert gao.robust_decode_batch(z, rows) == bads
        assert wb.robust_decode_batch(z, rows) == bads


def test_wb_config3_shape(rs):
    """BASELINE config 3 (n=64, t=21, 21 corrupted evaluations) on a few rows,
    against the C-speed property: decode returns the message and the exact
    error positions"""
    rng = random.Random(0xB203)
    n, t = 64, 21
    pt = _point(P, n, False)
    wb = rs.WelchBerlekampRobustDecoder(t, pt)
    gao = rs.GaoRobustDecoder(t, pt)
    rows, want = [], []
    for _ in range(6):
        c = [rng.randrange(P) for _ in range(t + 1)]
        enc = [orc.poly_eval(c, i + 1, P) for i in range(n)]
        bad = sorted(rng.sample(range(n), t))
        for i in bad:
            enc[i] = (enc[i] + 1 + rng.randrange(P - 1)) % P
        rows.append(enc)
        want.append((c, bad))
    assert wb.robust_decode_batch(list(range(n)), rows) == want
    assert gao.robust_decode_batch(list(range(n)), rows) == want


def test_config3_batch_property(rs):
    """BASELINE config 3 shape at a reduced batch (2048 words, n=64, t=21, exactly 21
    corrupted evaluations each): size-independent property on the limb boundary --
    both decoders return the message, and WB's result does not depend on the batch
    split."""
    from honeybadgermpc_b200 import ntl, robust

    n, t, batch = 64, 21, 2048
    k = t + 1
    xs = ntl.pack_vec(list(range(1, n + 1)), P)
    rng = np.random.default_rng(0xB203)
    msg = rng.integers(0, 2 ** 62, size=(batch, k, 4), dtype=np.uint64)
    enc = ntl.vandermonde_batch_evaluate_limbs(xs, msg, P)
    noise = rng.integers(0, 2 ** 62, size=(batch, n, 4), dtype=np.uint64)
    bad = np.zeros((batch, n), dtype=bool)
    for b in range(batch):
        bad[b, rng.choice(n, t, replace=False)] = True
    words = np.where(bad[:, :, None], noise, enc)
    coeffs, out_len, status = robust.wb_decode_batch_limbs(xs, words, k, (n - t) // 2, P)
    assert (status == 0).all() and np.array_equal(coeffs, msg)
    c2, _, s2 = robust.wb_decode_batch_limbs(xs, words[:100], k, (n - t) // 2, P)
    assert (s2 == 0).all() and np.array_equal(c2, msg[:100])
    gc, loc, ll, gst = robust.gao_decode_batch_limbs(xs, words, k, P)
    assert (gst == 0).all() and np.array_equal(gc, msg) and (ll == t + 1).all()
    # the locator vanishes exactly on the corrupted positions
    ev = ntl.vandermonde_batch_evaluate_limbs(xs, loc, P)
    assert np.array_equal(~ev.any(axis=2), bad)
