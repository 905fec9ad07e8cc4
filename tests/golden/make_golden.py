"""Generate tests/golden/golden_v1.json from the REFERENCE's own pure-Python
code (run in the authoring container, where /root/reference exists):

    python tests/golden/make_golden.py

Only reference code that never touches the (unbuildable) NTL extension is
recorded: ``polynomial.py`` (Polynomial.__call__/interpolate/evaluate_fft,
fnt_decode_step1/2, EvalPoint, get_omega) and ``reed_solomon_wb.py``.  These
compute the same maps as the NTL path (exact field arithmetic), so they pin
the oracle and the CUDA path.  All integers are stored as hex strings.
"""

import json
import os
import random
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def _hx(v):
    if v is None:
        return None
    if isinstance(v, (list, tuple)):
        return [_hx(w) for w in v]
    return hex(int(v))


def main():
    # a module that raises if the reference ever calls into "NTL"
    trap = types.ModuleType("trap")

    def _boom(*a, **k):
        raise RuntimeError("golden generation must not call the NTL layer")

    for name in [
        "lagrange_interpolate", "evaluate", "vandermonde_inverse",
        "vandermonde_batch_interpolate", "vandermonde_batch_evaluate", "fft", "partial_fft",
        "fft_batch_evaluate", "fft_interpolate", "fft_batch_interpolate", "SetNTLNumThreads",
        "AvailableNTLThreads", "gao_interpolate", "sqrt_mod", "SetNumThreads", "GetMaxThreads",
    ]:
        setattr(trap, name, _boom)
    trap.InterpolationError = type("InterpolationError", (Exception,), {})
    ref_shim.install(trap)

    from honeybadgermpc.field import GF
    from honeybadgermpc.polynomial import (
        EvalPoint, fnt_decode_step1, fnt_decode_step2, polynomials_over,
    )
    from honeybadgermpc.reed_solomon_wb import make_wb_encoder_decoder

    rng = random.Random(0xB200)
    out = {"modulus": hex(P), "generator": "tests/golden/make_golden.py"}
    fp = GF(P)
    poly = polynomials_over(fp)

    # A. EvalPoint ---------------------------------------------------------
    out["eval_points"] = []
    for n, use_omega in [(4, False), (4, True), (5, True), (16, True), (16, False),
                         (22, True), (64, True), (128, True)]:
        pt = EvalPoint(fp, n, use_omega_powers=use_omega)
        out["eval_points"].append({
            "n": n, "use_omega_powers": use_omega, "order": pt.order,
            "omega2": _hx(pt.omega2.value) if use_omega else None,
            "omega": _hx(pt.omega.value) if use_omega else None,
            "points": _hx([pt(i).value for i in range(n)]),
        })

    # B. encode = evaluate at the party points (Polynomial.__call__) --------
    out["encode"] = []
    for n, k, use_omega, batch in [(4, 2, False, 3), (4, 2, True, 3), (16, 6, True, 4),
                                   (16, 6, False, 2), (16, 16, False, 2), (64, 22, False, 1),
                                   (64, 22, True, 1), (128, 43, True, 1), (7, 3, True, 2)]:
        pt = EvalPoint(fp, n, use_omega_powers=use_omega)
        rows = [[rng.randrange(P) for _ in range(k)] for _ in range(batch)]
        enc = [[poly(r)(pt(i)).value for i in range(n)] for r in rows]
        out["encode"].append({"n": n, "k": k, "use_omega_powers": use_omega,
                              "coeffs": _hx(rows), "encoded": _hx(enc)})

    # C. interpolate from a subset (Polynomial.interpolate, Lagrange) ------
    out["interpolate"] = []
    for n, k, use_omega in [(4, 2, False), (4, 2, True), (16, 6, True), (16, 6, False),
                            (64, 22, True), (16, 16, False)]:
        pt = EvalPoint(fp, n, use_omega_powers=use_omega)
        zs = sorted(rng.sample(range(n), k))
        rng.shuffle(zs)
        ys = [rng.randrange(P) for _ in range(k)]
        poly._lagrange_cache.clear()
        f = poly.interpolate([(pt(z), fp(y)) for z, y in zip(zs, ys)])
        coeffs = [c.value for c in f.coeffs] + [0] * (k - len(f.coeffs))
        out["interpolate"].append({"n": n, "k": k, "use_omega_powers": use_omega,
                                   "zs": zs, "ys": _hx(ys), "coeffs": _hx(coeffs)})

    # D. python FFT (polynomial.py:271-302) -------------------------------
    out["fft"] = []
    for n, d in [(2, 2), (4, 3), (16, 6), (32, 20), (128, 43), (256, 256)]:
        pt = EvalPoint(fp, n, use_omega_powers=True)
        c = [rng.randrange(P) for _ in range(d)]
        ev = poly(c).evaluate_fft(pt.omega, n)
        out["fft"].append({"n": n, "omega": _hx(pt.omega.value), "coeffs": _hx(c),
                           "evals": _hx([e.value for e in ev])})

    # E. python fnt_decode (polynomial.py:305-382) ------------------------
    out["fnt_decode"] = []
    for n, k in [(8, 3), (16, 6), (32, 22), (128, 43)]:
        pt = EvalPoint(fp, n, use_omega_powers=True)
        zs = rng.sample(range(n), k)
        c = [rng.randrange(P) for _ in range(k)]
        ys = [poly(c)(pt(z)) for z in zs]
        as_, ais_ = fnt_decode_step1(poly, zs, pt.omega2, n)
        prec = fnt_decode_step2(poly, zs, ys, as_, ais_, pt.omega2, n)
        assert [v.value for v in prec.coeffs] == c
        out["fnt_decode"].append({"n": n, "omega": _hx(pt.omega.value), "zs": zs,
                                  "ys": _hx([y.value for y in ys]), "coeffs": _hx(c)})

    # F. Welch-Berlekamp (reed_solomon_wb.py) ------------------------------
    out["wb"] = []

    def wb_case(n, k, p, use_omega, msg, num_errors, num_nones, label):
        f = GF(p)
        pt = EvalPoint(f, n, use_omega_powers=use_omega) if (use_omega or p == P) else None
        enc, dec, _ = make_wb_encoder_decoder(n, k, p, pt)
        encoded = [v.value for v in enc(msg)]
        idx = rng.sample(range(n), num_errors + num_nones)
        recv = list(encoded)
        for i in idx[:num_errors]:
            v = rng.randrange(p)
            while v == encoded[i]:
                v = rng.randrange(p)
            recv[i] = v
        for i in idx[num_errors:]:
            recv[i] = None
        try:
            res = dec([None if v is None else f(v) for v in recv], debug=False)
            res = [c.value for c in res]
            err = None
        except Exception as e:  # noqa: BLE001
            res, err = None, f"{type(e).__name__}:{e}"
        out["wb"].append({"label": label, "n": n, "k": k, "p": hex(p),
                          "use_omega_powers": use_omega, "received": _hx(recv),
                          "decoded": _hx(res), "exception": err,
                          "error_positions": sorted(idx[:num_errors])})

    msg8 = [2, 3, 2, 8, 7, 5, 9, 5]
    wb_case(22, 8, 53, False, msg8, 0, 0, "p53 clean")
    wb_case(22, 8, 53, False, msg8, 0, 7, "p53 max erasures")
    wb_case(22, 8, 53, False, msg8, 3, 0, "p53 max errors")
    wb_case(22, 8, 53, False, msg8, 1, 1, "p53 mixed")
    wb_case(22, 8, 53, False, [0] * 8, 3, 0, "p53 zeros max errors")
    wb_case(22, 8, 53, False, [0] * 8, 1, 1, "p53 zeros mixed")
    wb_case(22, 8, 53, False, msg8, 8, 0, "p53 too many errors")
    wb_case(4, 2, P, False, [1, 2], 1, 0, "bls n4 one error")
    wb_case(4, 2, P, True, [1, 2], 1, 0, "bls n4 omega one error")
    wb_case(16, 6, P, False, [rng.randrange(P) for _ in range(6)], 5, 0, "bls n16 t5 max errors")
    wb_case(16, 6, P, True, [rng.randrange(P) for _ in range(6)], 3, 2, "bls n16 omega mixed")
    wb_case(16, 6, P, False, [rng.randrange(P) for _ in range(6)], 0, 0, "bls n16 clean")
    wb_case(16, 6, P, False, [rng.randrange(P) for _ in range(6)], 6, 0, "bls n16 too many errors")
    wb_case(64, 22, P, False, [rng.randrange(P) for _ in range(22)], 21, 0, "bls n64 t21 max errors")

    # G. open at zero (Polynomial.interpolate_at) ----------------------------
    out["interpolate_at_zero"] = []
    for n, t in [(4, 1), (16, 5)]:
        pt = EvalPoint(fp, n, use_omega_powers=False)
        c = [rng.randrange(P) for _ in range(t + 1)]
        shares = [(pt(i), poly(c)(pt(i))) for i in range(t + 1)]
        s = poly.interpolate_at(shares)
        assert s.value == c[0]
        out["interpolate_at_zero"].append({"n": n, "t": t,
                                           "shares": _hx([y.value for _, y in shares]),
                                           "secret": _hx(s.value)})

    path = os.path.join(HERE, "golden_v1.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=0, sort_keys=True)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
