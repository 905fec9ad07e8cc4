"""Generate tests/golden/wb_n64_v1.json: BASELINE configs[2] (n = 64, t = 21, Welch-Berlekamp
with t corrupted evaluations per word) decoded by the ORACLE's restatement of the reference's
pure-Python solver (reed_solomon_wb.py:79-151; ~1 s per word, which is why the vectors are
generated once here and committed) -- and, where /root/reference exists, cross-checked against
the reference's own ``WelchBerlekampRobustDecoder`` on the first words.

    python tests/golden/make_wb_n64_golden.py
"""

import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.dirname(os.path.dirname(HERE))]

from oracle import hbmpc_oracle as orc  # noqa: E402

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
N, T, WORDS = 64, 21, 36


def main():
    rng = random.Random(0xB203)
    out = {"generator": "tests/golden/make_wb_n64_golden.py", "n": N, "t": T, "cases": []}
    for omega in (False, True):
        pt = orc.EvalPoint(P, N, omega)
        xs = [pt(i) for i in range(N)]
        for w in range(WORDS // 2):
            msg = [rng.randrange(P) for _ in range(T + 1)]
            word = [orc.poly_eval(msg, x, P) for x in xs]
            n_err = T if w % 6 else rng.randrange(T)      # mostly the full t errors
            bad = sorted(rng.sample(range(N), n_err))
            for i in bad:
                word[i] = (word[i] + 1 + rng.randrange(P - 1)) % P
            coeffs, errors = orc.wb_robust_decode(list(range(N)), word, N, T + 1, P, pt)
            assert coeffs == msg and errors == bad
            gc, ge = orc.gao_robust_decode(list(range(N)), word, N, T + 1, P, pt)
            assert gc == msg and ge == bad
            out["cases"].append({"use_omega_powers": omega, "received": [hex(v) for v in word],
                                 "decoded": [hex(v) for v in coeffs], "errors": errors})
            print(len(out["cases"]), "words", flush=True)
    # the reference's own class on two words (the pure-Python solver is the reference's code)
    try:
        import ref_shim

        if ref_shim.reference_available():
            import logging

            logging.disable(logging.CRITICAL)
            ref_shim.install(orc)
            from honeybadgermpc.field import GF
            from honeybadgermpc.polynomial import EvalPoint
            from honeybadgermpc.reed_solomon import WelchBerlekampRobustDecoder

            for case in (out["cases"][0], out["cases"][WORDS // 2]):
                dec = WelchBerlekampRobustDecoder(T, EvalPoint(GF(P), N, case["use_omega_powers"]))
                got = dec.robust_decode(list(range(N)), [int(v, 16) for v in case["received"]])
                assert got == ([int(v, 16) for v in case["decoded"]], case["errors"])
            out["cross_checked_with_reference"] = 2
    except ImportError:
        pass
    with open(os.path.join(HERE, "wb_n64_v1.json"), "w") as fh:
        json.dump(out, fh, indent=0)
    print("wrote", len(out["cases"]), "cases")


if __name__ == "__main__":
    main()
