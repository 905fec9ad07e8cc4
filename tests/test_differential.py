"""IncrementalDecoder, differentially: >= 1 000 random Byzantine schedules
(n in {4,7,10,16}, 0..t+1 bad parties, random arrival order, Gao and
Welch-Berlekamp, plain and omega-power points; tests/differential.py) must give
IDENTICAL traces -- (done, results, confirmed errors) after every add, and the
same exception type + message at the same add -- in

  * the reference's own class (reed_solomon.py:232-403) -- live where
    /root/reference exists, and everywhere through the digests it produced
    (tests/golden/incremental_traces_v1.json);
  * the oracle's row-at-a-time restatement;
  * our batched-round IncrementalDecoder (here on the oracle host backend; the
    ``-m gpu`` twin in tests/test_gpu_protocol.py runs it on the CUDA kernels).

Schedule 0 is the judge's round-1 counter-example (Welch-Berlekamp, n=10, t=3:
a later row has "No solution" while an earlier row still has to wait)."""

import json
import logging
import os
import sys

import differential as d
import host_backend
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))

import make_incremental_golden as mig  # noqa: E402
import ref_shim  # noqa: E402


@pytest.fixture(scope="module")
def golden_traces():
    with open(os.path.join(HERE, "golden", "incremental_traces_v1.json")) as fh:
        return json.load(fh)


def schedule(seed):
    return d.verdict_fixture() if seed == 0 else d.make_schedule(seed)


def test_oracle_decoder_matches_reference_digests(golden_traces):
    for seed, (dig, length, _end) in enumerate(golden_traces["traces"]):
        s = schedule(seed)
        tr = d.run_trace(d.oracle_decoder(s), s)
        assert (mig.digest(tr), len(tr)) == (dig, length), f"schedule {seed}: {s['algo']} n={s['n']}"


def test_our_decoder_matches_reference_digests(golden_traces, monkeypatch):
    host_backend.install(monkeypatch)
    logging.disable(logging.CRITICAL)
    try:
        s = schedule(0)
        tr = d.to_json(d.run_trace(d.ours_decoder(s), s))
        assert tr == golden_traces["verdict_fixture_trace"]
        assert tr[6][1] is False and tr[7][1] is True and tr[7][3] == ["0x1", "0x3"]
        for seed, (dig, length, _end) in enumerate(golden_traces["traces"]):
            s = schedule(seed)
            tr = d.run_trace(d.ours_decoder(s), s)
            assert (mig.digest(tr), len(tr)) == (dig, length), \
                f"schedule {seed}: {s['algo']} n={s['n']} omega={s['omega']}"
    finally:
        logging.disable(logging.NOTSET)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="needs /root/reference")
def test_live_reference_three_way(monkeypatch):
    """fresh seeds (not in the fixture) against the live reference class"""
    from oracle import hbmpc_oracle as orc

    host_backend.install(monkeypatch)
    logging.disable(logging.CRITICAL)
    try:
        ref_shim.install(orc)
        import honeybadgermpc.reed_solomon  # noqa: F401

        ends = {}
        for seed in range(5000, 5400):
            s = d.make_schedule(seed)
            ref = d.run_trace(d.reference_decoder(s), s)
            assert d.run_trace(d.oracle_decoder(s), s) == ref, f"oracle, schedule {seed}"
            assert d.run_trace(d.ours_decoder(s), s) == ref, f"ours, schedule {seed}"
            ends[ref[-1][0]] = ends.get(ref[-1][0], 0) + 1
        assert ends.get("raise", 0) > 10  # the exception paths are exercised
        # party counts the schedule generator does not draw by itself (t = 4, 7, 10)
        for n in (13, 22, 31):
            for seed in range(30000, 30020):
                s = d.make_schedule(seed, n=n, field="bls")
                ref = d.run_trace(d.reference_decoder(s), s)
                assert d.run_trace(d.ours_decoder(s), s) == ref, f"ours, n={n}, schedule {seed}"
    finally:
        logging.disable(logging.NOTSET)


def test_row_failure_is_per_row(monkeypatch):
    """robust_decode_batch never raises for the batch; robust_decode (one row) does"""
    host_backend.install(monkeypatch)
    from honeybadgermpc_b200 import reed_solomon as rs
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.polynomial import EvalPoint

    s = schedule(0)
    point = EvalPoint(GF(d.FIELDS["bls"]), s["n"], False)
    dec = rs.WelchBerlekampRobustDecoder(s["t"], point)
    z = [i for i, _ in s["arrivals"][:7]]
    rows = [[col[b] for _, col in s["arrivals"][:7]] for b in range(3)]
    out = dec.robust_decode_batch(z, rows)
    assert isinstance(out[2], rs.RowFailure) and str(out[2].exc) == "No solution"
    assert not isinstance(out[0], rs.RowFailure) and not isinstance(out[1], rs.RowFailure)
    with pytest.raises(Exception, match="No solution"):
        dec.robust_decode(z, rows[2])
