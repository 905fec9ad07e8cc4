"""Shared body of the offline-caller parity tests (CPU: oracle host backend; GPU: kernels):
every function of honeybadgermpc_b200.offline against the oracle's restatement of what the
reference's caller computes with ``honeybadgermpc.ntl`` (offline_randousha.py,
progs/triple_refinement.py, preprocessing.py)."""

import random

import numpy as np
from conftest import BLS12_381_R as P

from oracle import hbmpc_oracle as orc


def check_offline_callers(batch=37, seed=5):
    from honeybadgermpc_b200 import offline
    from honeybadgermpc_b200.ntl import pack_rows, unpack_rows

    rng = random.Random(seed)
    for n, t in ((4, 1), (7, 2), (16, 5)):
        xs = list(range(1, n + 1))
        # --- randousha: share generation (:50-53), refinement (:72-78), check (:99-121)
        secrets = [rng.randrange(P) for _ in range(batch)]
        c_t = [[s] + [rng.randrange(P) for _ in range(t)] for s in secrets]
        c_2t = [[s] + [rng.randrange(P) for _ in range(2 * t)] for s in secrets]
        sh_t = offline.randousha_share(pack_rows(c_t, t + 1, P), n, P)
        sh_2t = offline.randousha_share(pack_rows(c_2t, 2 * t + 1, P), n, P)
        assert unpack_rows(sh_t) == orc.vandermonde_batch_evaluate(xs, c_t, P)
        assert unpack_rows(sh_2t) == orc.vandermonde_batch_evaluate(xs, c_2t, P)
        received = [[rng.randrange(P) for _ in range(n)] for _ in range(batch)]
        kept, chk = offline.randousha_refine(pack_rows(received, n, P), n, t, P)
        want = orc.vandermonde_batch_evaluate(xs, received, P)
        assert unpack_rows(kept) == [r[: n - 2 * t] for r in want]
        assert unpack_rows(chk) == [r[n - 2 * t:] for r in want]
        assert offline.randousha_check(sh_t, sh_2t, n, t, P) is True
        deg, sec = offline.degree_and_secret(sh_2t, n, P)
        assert deg.tolist() == [2 * t] * batch and unpack_rows(sec[:, None, :]) == [[s] for s in secrets]
        bad = sh_t.copy()
        bad[3, 0, 0] ^= np.uint64(1)                      # one wrong share: degree jumps to n-1
        assert offline.randousha_check(bad, sh_2t, n, t, P) is False
        other = offline.randousha_share(pack_rows([[s ^ 1] + r[1:] for s, r in zip(secrets, c_t)], t + 1, P), n, P)
        assert offline.randousha_check(other, sh_2t, n, t, P) is False   # secrets differ
        zero = np.zeros((2, n, 4), np.uint64)
        assert offline.degree_and_secret(zero, n, P)[0].tolist() == [0, 0]  # get_degree(0) = 0

        # --- refine_triples (triple_refinement.py:36-88) for m = n dirty triples per instance
        m = n
        d = (m - 1) // 2
        a = [[rng.randrange(P) for _ in range(m)] for _ in range(batch)]
        b = [[rng.randrange(P) for _ in range(m)] for _ in range(batch)]
        ac, bc, ar, br = offline.refine_triples_stage1(pack_rows(a, m, P), pack_rows(b, m, P), n, t, P)
        first, more = list(range(d + 1)), list(range(d + 1, 2 * d + 1))
        wa = orc.vandermonde_batch_interpolate(first, [r[: d + 1] for r in a], P)
        wb = orc.vandermonde_batch_interpolate(first, [r[: d + 1] for r in b], P)
        assert unpack_rows(ac) == wa and unpack_rows(bc) == wb
        assert unpack_rows(ar) == orc.vandermonde_batch_evaluate(more, wa, P)
        assert unpack_rows(br) == orc.vandermonde_batch_evaluate(more, wb, P)
        c_first = [[rng.randrange(P) for _ in range(d + 1)] for _ in range(batch)]
        c_rest = [[rng.randrange(P) for _ in range(d)] for _ in range(batch)]
        pv, qv, pq = offline.refine_triples_stage2(ac, bc, pack_rows(c_first, d + 1, P),
                                                   pack_rows(c_rest, d, P) if d else np.zeros((batch, 0, 4), np.uint64),
                                                   n, t, P)
        k = d + 1 - t
        fresh = list(range(n + 1, n + 1 + k))
        wc = orc.vandermonde_batch_interpolate(list(range(2 * d + 1)), [f + r for f, r in zip(c_first, c_rest)], P)
        assert unpack_rows(pv) == orc.vandermonde_batch_evaluate(fresh, wa, P)
        assert unpack_rows(qv) == orc.vandermonde_batch_evaluate(fresh, wb, P)
        assert unpack_rows(pq) == orc.vandermonde_batch_evaluate(fresh, wc, P)
        # honest triples stay triples: with c = a*b on all 2d+1 points, pq = p*q
        fa = [[orc.poly_eval(w, x, P) for x in range(2 * d + 1)] for w in wa]
        fb = [[orc.poly_eval(w, x, P) for x in range(2 * d + 1)] for w in wb]
        prod = [[(u * v) % P for u, v in zip(ra, rb)] for ra, rb in zip(fa, fb)]
        _, _, pq2 = offline.refine_triples_stage2(
            ac, bc, pack_rows([r[: d + 1] for r in prod], d + 1, P),
            pack_rows([r[d + 1:] for r in prod], d, P) if d else np.zeros((batch, 0, 4), np.uint64), n, t, P)
        assert unpack_rows(pq2) == [[(u * v) % P for u, v in zip(ru, rv)]
                                    for ru, rv in zip(unpack_rows(pv), unpack_rows(qv))]

        # --- preprocessing._write_polys (:222-231)
        polys = [[rng.randrange(P) for _ in range(t + 1)] for _ in range(batch)]
        vals = offline.write_polys_values(pack_rows(polys, t + 1, P), n, P)
        ev = orc.vandermonde_batch_evaluate(xs, polys, P)
        assert unpack_rows(vals) == [[row[i] for row in ev] for i in range(n)]
