"""The codec objects, IncrementalDecoder and batch_reconstruct on the CUDA
path: the reference's known answers (tests/test_reed_solomon.py,
tests/test_batch_reconstruction.py, tests/test_offline_randousha.py algebra)
and randomized parity against the oracle.  ``pytest -m gpu``."""

import asyncio
import random

import pytest
from conftest import BLS12_381_R as P
from sim_net import SimNet, run

from oracle import hbmpc_oracle as orc

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def rs():
    from honeybadgermpc_b200 import ntl, reed_solomon

    ntl._ctx(P)
    return reed_solomon


def _point(n, omega, p=P):
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.polynomial import EvalPoint

    return EvalPoint(GF(p), n, omega)


# --- tests/test_reed_solomon.py:19-165 ---------------------------------------


def test_encoder_decoder_kats(rs):
    pt = _point(4, False)
    for enc in (rs.VandermondeEncoder(pt), rs.EncoderFactory.get(pt),
                rs.EncoderFactory.get(pt, rs.Algorithm.VANDERMONDE)):
        assert enc.encode([1, 2]) == [3, 5, 7, 9]
        assert enc.encode([[1, 2], [2, 3]]) == [[3, 5, 7, 9], [5, 8, 11, 14]]
    for dec in (rs.VandermondeDecoder(pt), rs.DecoderFactory.get(pt)):
        assert dec.decode([1, 3], [5, 9]) == [1, 2]
        assert dec.decode([1, 3], [[5, 9], [8, 14]]) == [[1, 2], [2, 3]]
    pto = _point(4, True)
    w = pto.omega.value
    want = [(1 + 2 * pow(w, i, P)) % P for i in range(4)]
    want2 = [(2 + 3 * pow(w, i, P)) % P for i in range(4)]
    for enc in (rs.FFTEncoder(pto), rs.VandermondeEncoder(pto), rs.OptimalEncoder(pto),
                rs.EncoderFactory.get(pto), rs.EncoderFactory.get(pto, rs.Algorithm.FFT)):
        assert enc.encode([1, 2]) == want
        assert enc.encode([[1, 2], [2, 3]]) == [want, want2]
    for dec in (rs.FFTDecoder(pto), rs.VandermondeDecoder(pto), rs.OptimalDecoder(pto)):
        assert dec.decode([1, 3], [want[1], want[3]]) == [1, 2]
        assert dec.decode([1, 3], [[want[1], want[3]], [want2[1], want2[3]]]) == [[1, 2], [2, 3]]
    with pytest.raises(ValueError):
        rs.EncoderFactory.get(pt, "nope")
    with pytest.raises(ValueError):
        rs.DecoderFactory.get(pt, rs.Algorithm.GAO)
    with pytest.raises(ValueError):
        rs.RobustDecoderFactory.get(1, pt, rs.Algorithm.FFT)


def test_selectors(rs, monkeypatch):
    """tests/test_reed_solomon.py:186-277: which class the heuristics pick"""
    from honeybadgermpc_b200 import ntl

    monkeypatch.setattr("psutil.cpu_count", lambda logical=False: 4)
    for n, cls in [(4, rs.VandermondeEncoder), (8, rs.FFTEncoder), (9, rs.VandermondeEncoder),
                   (13, rs.FFTEncoder), (65, rs.VandermondeEncoder), (100, rs.FFTEncoder),
                   (128, rs.FFTEncoder), (129, rs.FFTEncoder)]:
        assert type(rs.EncoderSelector.select(_point(n, True), 1)) is cls, n
    rs.DecoderSelector.set_optimal_thread_count(100)
    assert ntl.AvailableNTLThreads() == 4
    assert type(rs.DecoderSelector.select(_point(4, True), 1)) is rs.VandermondeDecoder
    nt = ntl.AvailableNTLThreads()
    assert type(rs.DecoderSelector.select(_point(16, True), int(0.5 * 16 * nt))) is rs.FFTDecoder
    assert type(rs.DecoderSelector.select(_point(16, True), int(0.5 * 16 * nt) + 1)) is rs.VandermondeDecoder
    rs.DecoderSelector.set_optimal_thread_count(1)
    assert ntl.AvailableNTLThreads() == 1


@pytest.mark.parametrize("n,t,omega", [(4, 1, False), (7, 2, True), (16, 5, False), (16, 5, True)])
def test_codec_round_trip_vs_oracle(rs, n, t, omega):
    rng = random.Random(n * 10 + t)
    pt = _point(n, omega)
    opt = orc.EvalPoint(P, n, omega)
    xs = [opt(i) for i in range(n)]
    enc, dec = rs.EncoderFactory.get(pt), rs.DecoderFactory.get(pt)
    polys = [[rng.randrange(P) for _ in range(t + 1)] for _ in range(33)]
    encoded = enc.encode(polys)
    assert encoded == orc.vandermonde_batch_evaluate(xs, polys, P)
    z = rng.sample(range(n), t + 1)
    assert dec.decode(z, [[row[i] for i in z] for row in encoded]) == polys


# --- IncrementalDecoder -------------------------------------------------------


def _inc(rs, n, t, omega, batch, algo="gao"):
    pt = _point(n, omega)
    kind = rs.Algorithm.FFT if omega else rs.Algorithm.VANDERMONDE
    return pt, rs.IncrementalDecoder(
        rs.EncoderFactory.get(pt, kind), rs.DecoderFactory.get(pt, kind),
        rs.RobustDecoderFactory.get(t, pt, algo), degree=t, batch_size=batch, max_errors=t)


@pytest.mark.parametrize("omega", [False, True])
@pytest.mark.parametrize("algo", ["gao", "welch-berlekamp"])
def test_incremental_decoder(rs, omega, algo):
    rng = random.Random(17)
    n, t, batch = 7, 2, 9
    opt = orc.EvalPoint(P, n, omega)
    polys = [[rng.randrange(P) for _ in range(t + 1)] for _ in range(batch)]
    cols = [[orc.poly_eval(c, opt(i), P) for c in polys] for i in range(n)]

    # no faults: done after t+1+t columns, in any arrival order
    pt, inc = _inc(rs, n, t, omega, batch, algo)
    order = [4, 0, 6, 2, 5, 1, 3]
    for count, i in enumerate(order, 1):
        inc.add(i, cols[i])
        assert inc.done() == (count >= 2 * t + 1)
        if inc.done():
            break
    res, errs = inc.get_results()
    assert res == polys and errs == set()
    inc.add(3, cols[3])  # ignored once done

    # duplicate columns are ignored, wrong length is rejected
    pt, inc = _inc(rs, n, t, omega, batch, algo)
    inc.add(0, cols[0])
    inc.add(0, cols[1])
    with pytest.raises(rs.DecodeValidationError):
        inc.add(1, cols[1][:-1])
    assert inc.get_results() == (None, None)

    if algo == "gao":
        # one Byzantine party inside the first t+1 columns, another one later
        bad = {1: [(v + 1) % P for v in cols[1]], 5: [0] * batch}
        pt, inc = _inc(rs, n, t, omega, batch, algo)
        for i in [1, 0, 2, 3, 5, 4]:
            inc.add(i, bad.get(i, cols[i]))
            assert not inc.done()
        inc.add(6, cols[6])
        assert inc.done()
        res, errs = inc.get_results()
        assert res == polys and errs == {1, 5}
    else:
        # one Byzantine party (beyond-capacity words, where the reference's solver raises
        # "No solution" uncaught, are covered by test_incremental_decoder_differential)
        bad = {4: [0] * batch}
        pt, inc = _inc(rs, n, t, omega, batch, algo)
        for i in [0, 1, 2, 3, 4]:
            inc.add(i, bad.get(i, cols[i]))
            assert not inc.done()
        inc.add(5, cols[5])
        assert inc.done()
        res, errs = inc.get_results()
        assert res == polys and errs == {4}

    # a party that lies in ONE row only is still found and evicted
    sly = list(cols[2])
    sly[4] = (sly[4] + 5) % P
    pt, inc = _inc(rs, n, t, omega, batch, algo)
    for i in [0, 1, 2, 3, 4, 5]:
        inc.add(i, sly if i == 2 else cols[i])
    assert inc.done()
    res, errs = inc.get_results()
    assert res == polys and errs == {2}


def test_incremental_decoder_differential(rs):
    """>= 1 000 random Byzantine schedules (tests/differential.py) on the CUDA kernels:
    the trace of our batched-round decoder must equal the trace of the oracle's
    row-at-a-time restatement of reed_solomon.py:334-365 AND the digest the reference's own
    class produced for the same schedule (tests/golden/incremental_traces_v1.json).
    Schedule 0 is the round-1 Welch-Berlekamp counter-example."""
    import json
    import logging
    import os
    import sys

    import differential as d

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_incremental_golden as mig

    with open(os.path.join(here, "golden", "incremental_traces_v1.json")) as fh:
        gold = json.load(fh)
    logging.disable(logging.CRITICAL)
    try:
        ends = {}
        for seed, (dig, length, end) in enumerate(gold["traces"]):
            s = d.verdict_fixture() if seed == 0 else d.make_schedule(seed)
            ours = d.run_trace(d.ours_decoder(s), s)
            if seed == 0:
                assert d.to_json(ours) == gold["verdict_fixture_trace"]
            if seed % 4 == 0:  # the pure-Python oracle is the slow side
                assert ours == d.run_trace(d.oracle_decoder(s), s), f"schedule {seed} vs oracle"
            assert (mig.digest(ours), len(ours)) == (dig, length), \
                f"schedule {seed}: {s['algo']} n={s['n']} omega={s['omega']} {s['field']}"
            ends[end] = ends.get(end, 0) + 1
        assert ends["Exception"] > 50 and ends["AssertionError"] > 10 and ends["done"] > 500
    finally:
        logging.disable(logging.NOTSET)


def test_incremental_decoder_device_resident(rs, monkeypatch):
    """The device-resident column path (hbg_columns_to_rows / hbg_interpolate_reencode /
    hbg_compare_columns) must be indistinguishable from the host path and from the reference:
    (a) the differential schedules again with every decoder forced onto the device path,
    (b) a batch large enough for the tensor-core kernel, honest and with one lying party."""
    import json
    import logging
    import os
    import sys

    import differential as d
    import numpy as np

    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.join(here, "golden"))
    import make_incremental_golden as mig

    with open(os.path.join(here, "golden", "incremental_traces_v1.json")) as fh:
        gold = json.load(fh)
    monkeypatch.setattr(rs, "DEVICE_MIN_BATCH", 1)
    before = rs._DeviceColumns.totals["decoders"]
    logging.disable(logging.CRITICAL)
    try:
        for seed in range(0, 400):
            s = d.verdict_fixture() if seed == 0 else d.make_schedule(seed)
            ours = d.run_trace(d.ours_decoder(s), s)
            dig, length, _ = gold["traces"][seed]
            assert (mig.digest(ours), len(ours)) == (dig, length), f"schedule {seed} on the device path"
        assert rs._DeviceColumns.totals["decoders"] - before == 400

        n, t, batch = 16, 5, 3000
        rng = random.Random(77)
        for omega in (False, True):
            pt = _point(n, omega)
            polys = [[rng.randrange(P) for _ in range(t + 1)] for _ in range(batch)]
            cols = np.ascontiguousarray(
                rs.EncoderFactory.get(pt).encode_batch_limbs(rs.pack_rows(polys, t + 1, P)).transpose(1, 0, 2))
            lie = cols[7].copy()
            lie[batch - 1, 0] ^= np.uint64(1)  # one element of one row
            for bad in (None, 7):
                traces = []
                for device in (True, False):
                    inc = rs.IncrementalDecoder(rs.EncoderFactory.get(pt), rs.DecoderFactory.get(pt),
                                                rs.RobustDecoderFactory.get(t, pt), t, batch, t, device=device)
                    tr = []
                    for i in [3, 9, 0, 7, 12, 15, 1, 2, 4, 5, 6, 8, 10]:
                        inc.add(i, (lie if i == bad else cols[i]).tobytes())
                        tr.append((inc.done(), sorted(inc.get_results()[1] or [])))
                        if inc.done():
                            break
                    res, errs = inc.get_results_limbs()
                    assert np.array_equal(res, rs.pack_rows(polys, t + 1, P)) and errs == ({bad} if bad else set())
                    traces.append(tr)
                assert traces[0] == traces[1]
                assert len(traces[0]) == (11 if bad is None else 12)
    finally:
        logging.disable(logging.NOTSET)


def test_offline_callers_on_limb_path(rs):
    """offline_randousha.py:47-53,72-78,99-121; progs/triple_refinement.py:36-88;
    preprocessing.py:222-231 -- the compute steps on the kernels vs the oracle"""
    import offline_cases

    offline_cases.check_offline_callers(batch=37)
    offline_cases.check_offline_callers(batch=700, seed=9)  # large enough for the tensor-core kernel


# --- batch_reconstruct (tests/test_batch_reconstruction.py:12-170) -------------


async def _reconstruct(n, t, shares, omega, skip=(), zero=(), delay=0.0, config=None, wire="ints"):
    from honeybadgermpc_b200.batch_reconstruction import batch_reconstruct
    from honeybadgermpc_b200.field import GF

    fp = GF(P)
    net = SimNet(n, max_delay=delay, seed=5)
    jobs = []
    for i in range(n):
        if i in skip:
            continue
        ss = [fp(0) for _ in shares[i]] if i in zero else [fp(v) for v in shares[i]]
        jobs.append(batch_reconstruct(ss, P, t, n, i, net.sends[i], net.recvs[i],
                                      use_omega_powers=omega, config=config, wire=wire))
    return await asyncio.gather(*jobs)


def test_batch_reconstruct_kats():
    from honeybadgermpc_b200.field import GFElement

    shares = [(3, 7, 4), (4, 10, 6), (5, 13, 8), (6, 16, 10)]  # x+2, 3x+4, 2x+2
    for zero in ((), (1,)):
        for wire in ("ints", "limbs"):
            for r in run(_reconstruct(4, 1, shares, False, zero=zero, delay=0.01, wire=wire)):
                assert all(type(e) is GFElement for e in r)
                assert r == [2, 4, 2]
    w = _point(4, True).omega.value
    oshares = [((pow(w, i, P) + 2) % P, (3 * pow(w, i, P) + 4) % P) for i in range(4)]
    for zero in ((), (1,)):
        for r in run(_reconstruct(4, 1, oshares, True, zero=zero)):
            assert r == [2, 4]


@pytest.mark.parametrize("n,t,count,omega,algo", [(4, 1, 256, False, "gao"), (7, 2, 100, True, "gao"),
                                                  (16, 5, 64, False, "welch-berlekamp"),
                                                  (16, 5, 601, True, "gao")])
def test_batch_reconstruct_random(n, t, count, omega, algo):
    """BASELINE config 1 shape (n=4, t=1, 256 shares) and larger: every honest
    party opens the secrets, also with t parties sending zeros"""
    rng = random.Random(count)
    opt = orc.EvalPoint(P, n, omega)
    secrets = [rng.randrange(P) for _ in range(count)]
    polys = [[s] + [rng.randrange(P) for _ in range(t)] for s in secrets]
    shares = [[orc.poly_eval(c, opt(i), P) for c in polys] for i in range(n)]

    class Cfg:
        induce_faults = False
        decoding_algorithm = algo

    nbad = t if algo == "gao" else 1  # WB raises beyond capacity, see test_incremental_decoder
    for zero in ((), tuple(rng.sample(range(n), nbad))):
        for wire in ("ints", "limbs"):
            results = run(_reconstruct(n, t, shares, omega, zero=zero, delay=0.002, config=Cfg, wire=wire))
            for r in results:
                assert [e.value for e in r] == secrets


def test_randousha_refinement_algebra(rs):
    """offline_randousha.py:72-78: the hyper-invertible step is
    encoder.encode(transpose(received)) on the points 1..n (config 4 shape)"""
    rng = random.Random(4)
    n, t, rows = 16, 5, 50
    pt = _point(n, False)
    enc = rs.EncoderFactory.get(pt)
    received = [[rng.randrange(P) for _ in range(n)] for _ in range(rows)]
    out = enc.encode(received)
    xs = list(range(1, n + 1))
    assert out == orc.vandermonde_batch_evaluate(xs, received, P)
    # degree check of the reference (offline_randousha.py:105-110): decode from all n points
    dec = rs.DecoderFactory.get(pt)
    polys = [[rng.randrange(P) for _ in range(t + 1)] + [0] * (n - t - 1) for _ in range(rows)]
    shares = enc.encode(polys)
    assert dec.decode(list(range(n)), shares) == polys


@pytest.mark.parametrize("omega", [False, True])
def test_robust_reconstruct_single_share(rs, omega):
    """robust_reconstruction.py:14-30 -- one shared value, one faulty party"""
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.robust_reconstruction import robust_reconstruct

    n, t = 7, 2
    fp = GF(P)
    pt = _point(n, omega)
    rng = random.Random(3)
    coeffs = [rng.randrange(P) for _ in range(t + 1)]
    shares = [orc.poly_eval(coeffs, pt(i).value, P) for i in range(n)]
    shares[5] = (shares[5] + 1) % P  # party 5 is the second to arrive (delays below)

    async def go():
        loop = asyncio.get_event_loop()
        futs = []
        for i in range(n):
            f = loop.create_future()
            loop.call_later(0.001 * ((i * 3) % n), f.set_result, fp(shares[i]))
            futs.append(f)
        return await robust_reconstruct(futs, fp, n, t, pt, t)

    poly, errors = run(go())
    assert poly == coeffs and errors == {5}


# --- n parties on n GPUs (party_sim.py): the CUDA codec, all parties in one process ---


def _party_shares(n, t, batch, omega, seed):
    import numpy as np
    import torch

    rng = random.Random(seed)
    pt = _point(n, omega)
    xs = [pt(i).value for i in range(n)]
    secrets = [rng.randrange(P) for _ in range(batch)]
    polys = [[s] + [rng.randrange(P) for _ in range(t)] for s in secrets]
    out = []
    for x in xs:
        vals = [sum(c * pow(x, e, P) for e, c in enumerate(f)) % P for f in polys]
        raw = b"".join(v.to_bytes(32, "little") for v in vals)
        out.append(torch.from_numpy(np.frombuffer(raw, dtype=np.int64).reshape(batch, 4).copy()).cuda())
    return secrets, out


def _limbs_to_ints(t):
    raw = t.cpu().contiguous().numpy().tobytes()
    return [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]


@pytest.mark.parametrize("n,t,batch,omega", [(4, 1, 256, False), (4, 1, 7, True), (16, 5, 1000, False),
                                              (16, 5, 65, True), (7, 2, 33, False), (2, 0, 5, False)])
def test_party_simulation_in_process(n, t, batch, omega):
    import torch

    from honeybadgermpc_b200 import party_sim

    secrets, shares = _party_shares(n, t, batch, omega, seed=n * 1000 + batch)
    codecs = [party_sim.CudaCodec(P, n, use_omega_powers=omega)] * n  # one GPU plays every party
    for got, ok in party_sim.simulate_in_process(codecs, shares, t):
        assert ok and _limbs_to_ints(got) == secrets
    # a wrong share of the last party is noticed by everyone (re-encode + compare) and, with
    # t >= 1, corrected by the robust fallback (hbg_gao_decode_batch on device pointers)
    if n > t + 1:
        shares[n - 1] = shares[n - 1].clone()
        shares[n - 1][0, 0] += 1
        info = []
        res = party_sim.simulate_in_process(codecs, shares, t, info=info)
        if t == 0:
            assert not any(ok for _, ok in res)
        else:
            for got, ok in res:
                assert ok and _limbs_to_ints(got) == secrets
            assert all(errs == [n - 1] and rounds == 1 for errs, rounds in info)
            # t parties sending noise in both rounds
            liars = tuple(range(1, 1 + t)) if n - 1 not in range(1, 1 + t) else tuple(range(t))
            shares[n - 1][0, 0] -= 1
            info = []
            for got, ok in party_sim.simulate_in_process(codecs, shares, t, byzantine=liars, info=info):
                assert ok and _limbs_to_ints(got) == secrets
            assert all(errs == sorted(liars) and rounds == 2 for errs, rounds in info)
    torch.cuda.synchronize()


def _party_rank(rank, world, port, t, batch, results):
    import os

    import torch
    import torch.distributed as dist

    from honeybadgermpc_b200 import party_sim

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        secrets, shares = _party_shares(world, t, batch, False, seed=world + batch)
        got, ok = party_sim.batch_reconstruct_collective(shares[rank], t, party_sim.CudaCodec(P, world))
        results[rank] = bool(ok) and _limbs_to_ints(got) == secrets
    finally:
        dist.destroy_process_group()


def test_party_simulation_nccl():
    """one party per GPU over NCCL (needs >= 2 GPUs; the round-end single-GPU run skips it)"""
    import socket

    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    t = (world - 1) // 3
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_party_rank, args=(r, world, port, t, 1000, results)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert dict(results) == {r: True for r in range(world)}
