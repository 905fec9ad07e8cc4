"""The message pattern of the n-parties-on-n-GPUs simulation
(honeybadgermpc_b200/party_sim.py) on CPU: world_size 4 and 2, gloo backend,
with the Python oracle as the codec (the CUDA codec is tested in
tests/test_gpu_protocol.py)."""

import os
import random
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import BLS12_381_R as P
from conftest import ROOT

sys.path.insert(0, ROOT)
from honeybadgermpc_b200 import party_sim  # noqa: E402

from oracle import hbmpc_oracle as orc  # noqa: E402


def to_limbs(rows):
    """list of lists of ints -> int64 tensor [rows, width, 4]"""
    raw = b"".join(int(v).to_bytes(32, "little") for r in rows for v in r)
    a = np.frombuffer(raw, dtype=np.int64).reshape(len(rows), len(rows[0]) if rows else 0, 4)
    return torch.from_numpy(a.copy())


def to_ints(t):
    raw = t.contiguous().numpy().tobytes()
    flat = [int.from_bytes(raw[i:i + 32], "little") for i in range(0, len(raw), 32)]
    w = t.shape[1]
    return [flat[i * w:(i + 1) * w] for i in range(t.shape[0])]


class _OracleRobust:
    """robust_decode of the codec interface (Gao over all n parties) by the oracle"""

    def _points(self):
        raise NotImplementedError

    def robust_decode(self, rows, k):
        xs = self._points()
        coeffs, decoded, bad = [], [], []
        for word in to_ints(rows):
            dec, loc = orc.gao_interpolate(xs, word, k, self.p)
            decoded.append(dec is not None)
            coeffs.append(dec if dec is not None else [0] * k)
            roots = [len(loc or []) > 1 and orc.poly_eval(loc, x, self.p) == 0 for x in xs]
            bad.append(roots)
        return to_limbs(coeffs), torch.tensor(decoded), torch.tensor(bad).reshape(len(coeffs), self.n)


class OracleCodec(_OracleRobust):
    """the codec interface of party_sim on CPU tensors, computed by the oracle"""

    def _points(self):
        return self.xs

    def __init__(self, p, n):
        self.p, self.n = p, n
        self.xs = [i + 1 for i in range(n)]

    def encode(self, coeffs):
        return to_limbs(orc.vandermonde_batch_evaluate(self.xs, to_ints(coeffs), self.p))

    def interpolate(self, z, ys):
        return to_limbs(orc.vandermonde_batch_interpolate([self.xs[i] for i in z], to_ints(ys), self.p))


def make_shares(n, t, batch, seed):
    """secrets and every party's shares (party i: f_b(i + 1)), deterministic"""
    rng = random.Random(seed)
    secrets = [rng.randrange(P) for _ in range(batch)]
    polys = [[s] + [rng.randrange(P) for _ in range(t)] for s in secrets]
    shares = [[sum(c * pow(i + 1, e, P) for e, c in enumerate(f)) % P for f in polys] for i in range(n)]
    return secrets, shares


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, t, batch, corrupt, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        secrets, shares = make_shares(world, t, batch, seed=world * 100 + batch)
        mine = to_limbs([[v] for v in shares[rank]]).reshape(batch, 4)
        if corrupt and rank == world - 1:
            mine[0, 0] += 1  # a wrong share: <= t faulty parties are corrected by the robust fallback
        info = {}
        got, ok = party_sim.batch_reconstruct_collective(mine, t, OracleCodec(P, world), info=info)
        opened = [r[0] for r in to_ints(got.reshape(batch, 1, 4))]
        results[rank] = (opened == secrets, ok, info["errors"])
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,t,batch,corrupt", [(4, 1, 7, False), (4, 1, 8, False), (2, 0, 5, False),
                                                   (4, 1, 6, True)])
def test_parties_as_ranks_gloo(world, t, batch, corrupt):
    port = _free_port()
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, t, batch, corrupt, results))
             for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(180)
        assert p.exitcode == 0
    res = dict(results)
    assert set(res) == set(range(world))
    assert all(match and ok for match, ok, _ in res.values())
    for r, (_, _, errors) in res.items():
        # R1 carries the lie (party world-1's chunk polynomial is off in one coefficient); by R2
        # every honest party publishes corrected values, so only R1 needs the robust decoder
        assert errors == ([world - 1] if corrupt else [])


class OracleOmegaCodec(_OracleRobust):
    """omega-power points (FFT encode / interpolate of the oracle)"""

    def _points(self):
        return [self.pt(i) for i in range(self.n)]

    def __init__(self, p, n):
        self.p, self.n = p, n
        self.pt = orc.EvalPoint(p, n, True)

    def encode(self, coeffs):
        return to_limbs(orc.fft_batch_evaluate(to_ints(coeffs), self.pt.omega, self.p, self.pt.order, self.n))

    def interpolate(self, z, ys):
        return to_limbs(orc.fft_batch_interpolate(list(z), to_ints(ys), self.pt.omega, self.p, self.pt.order))


@pytest.mark.parametrize("n,t,batch", [(16, 5, 20), (7, 2, 5)])
def test_in_process_simulation_omega_points(n, t, batch):
    rng = random.Random(n)
    codec = OracleOmegaCodec(P, n)
    xs = [codec.pt(i) for i in range(n)]
    secrets = [rng.randrange(P) for _ in range(batch)]
    polys = [[s] + [rng.randrange(P) for _ in range(t)] for s in secrets]
    per_party = [to_limbs([[sum(c * pow(x, e, P) for e, c in enumerate(f)) % P] for f in polys]).reshape(batch, 4)
                 for x in xs]
    for got, ok in party_sim.simulate_in_process([codec] * n, per_party, t):
        assert ok and [r[0] for r in to_ints(got.reshape(batch, 1, 4))] == secrets
    per_party[n - 1][0, 0] += 1
    info = []
    for got, ok in party_sim.simulate_in_process([codec] * n, per_party, t, info=info):
        assert ok and [r[0] for r in to_ints(got.reshape(batch, 1, 4))] == secrets
    assert all(errs == [n - 1] and rounds == 1 for errs, rounds in info)
    # a party that sends noise in both rounds
    info = []
    for j, (got, ok) in enumerate(party_sim.simulate_in_process([codec] * n, per_party, t, byzantine=(2,), info=info)):
        assert ok and [r[0] for r in to_ints(got.reshape(batch, 1, 4))] == secrets
    assert all(errs == [2, n - 1] and rounds == 2 for errs, rounds in info)


def test_in_process_simulation_matches():
    n, t, batch = 4, 1, 9
    secrets, shares = make_shares(n, t, batch, seed=3)
    codecs = [OracleCodec(P, n) for _ in range(n)]
    per_party = [to_limbs([[v] for v in shares[i]]).reshape(batch, 4) for i in range(n)]
    for got, ok in party_sim.simulate_in_process(codecs, per_party, t):
        assert ok and [r[0] for r in to_ints(got.reshape(batch, 1, 4))] == secrets
