#!/usr/bin/env python
"""Secondary measurements: the other BASELINE.json configs on ONE B200 (the
headline config is bench.py).  Prints one JSON object per config.

    python tools/bench_configs.py [--quick]
"""
import argparse
import asyncio
import json
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from honeybadgermpc_b200 import _native, ntl, robust  # noqa: E402
from honeybadgermpc_b200.field import GF  # noqa: E402
from honeybadgermpc_b200.polynomial import EvalPoint  # noqa: E402

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
E = 32


def synth(batch, width, seed):
    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2 ** 63, size=(batch, width, 4), dtype=np.uint64)
    a[:, :, 3] >>= np.uint64(2)
    return a


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).cuda()


def timed(fn, reps, stream):
    for _ in range(3):
        fn()
    stream.synchronize()
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for _ in range(reps):
        fn()
    t1.record(stream)
    stream.synchronize()
    return t0.elapsed_time(t1) / reps


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    args = ap.parse_args()
    ctx = _native.get_context(P, 0)
    stream = torch.cuda.Stream()
    ctx.set_stream(stream.cuda_stream)
    out = []
    with torch.cuda.stream(stream):
        # ---- config 4: RanDouSha refinement = Vandermonde encode of 16 received shares on 1..16
        n, rows = 16, 5462
        xs = ntl.pack_vec(list(range(1, n + 1)), P)
        rec = [dev(synth(rows, n, 40 + i)) for i in range(8)]
        res = torch.empty((rows, n, 4), dtype=torch.int64, device="cuda")
        it = [0]

        def cfg4():
            ctx.vandermonde_batch_evaluate(xs, rec[it[0] % 8].data_ptr(), rows, n, res.data_ptr(), _native.MEM_DEVICE)
            it[0] += 1

        ms = timed(cfg4, 50, stream)
        out.append({"config": "cfg4 RanDouSha refinement n=16 t=5 (one of the two 5462x16 Vandermonde encodes)",
                    "kernel": ctx.last_kernel(), "ms": ms, "outputs_per_s": rows * 6 / (ms * 1e-3),
                    "GBps_algorithmic": rows * 2 * n * E / (ms * 1e-3) / 1e9})

        # ---- config 5, one GPU's shard: n=128, t=42, NTT-128 encode + interpolate from 43 points
        n, k, batch = 128, 43, 131072 if not args.quick else 16384
        pt = EvalPoint(GF(P), n, True)
        omega = ntl.pack_vec([pt.omega.value], P)[0]
        c = dev(synth(batch, k, 50))
        e = torch.empty((batch, n, 4), dtype=torch.int64, device="cuda")
        ms_enc = timed(lambda: ctx.fft_batch_evaluate(omega, pt.order, c.data_ptr(), batch, k, n, e.data_ptr(),
                                                      _native.MEM_DEVICE), 5, stream)
        k_enc = ctx.last_kernel()
        enc_variants = {}
        for name, fft_path, mv_path in (("ntt_smem_kernel (IMAD butterflies)", "ntt", "auto"),
                                        ("tc_apply_kernel, streamed 128x43 DFT matrix", "matrix", "tc")):
            ctx.set_fft_path(fft_path)
            ctx.set_matvec_path(mv_path)
            e2 = torch.empty_like(e)
            enc_variants[name] = timed(lambda: ctx.fft_batch_evaluate(omega, pt.order, c.data_ptr(), batch, k, n,
                                                                      e2.data_ptr(), _native.MEM_DEVICE), 5, stream)
            stream.synchronize()
            assert torch.equal(e2, e), name
            del e2
        ctx.set_fft_path("auto")
        ctx.set_matvec_path("auto")
        zs = sorted(random.Random(5).sample(range(n), k))
        y = e.index_select(1, torch.tensor(zs, device="cuda")).contiguous()
        r = torch.empty((batch, k, 4), dtype=torch.int64, device="cuda")
        ms_dec = timed(lambda: ctx.fft_batch_interpolate(omega, pt.order, zs, y.data_ptr(), batch, r.data_ptr(),
                                                         _native.MEM_DEVICE), 5, stream)
        k_dec = ctx.last_kernel()
        stream.synchronize()
        assert torch.equal(r, c), "cfg5 round trip"
        out.append({"config": f"cfg5 shard n=128 t=42 batch={batch} (1/8 of 1 Mi)", "encode_kernel": k_enc,
                    "interpolate_kernel": k_dec, "encode_ms": ms_enc, "interpolate_ms": ms_dec,
                    "encode_ms_by_kernel": enc_variants,
                    "shares_per_s": batch * k / ((ms_enc + ms_dec) * 1e-3),
                    "GBps_algorithmic": batch * (3 * k + n) * E / ((ms_enc + ms_dec) * 1e-3) / 1e9})

        # ---- config 3: n=64, t=21, 21 corrupted evaluations per word, Welch-Berlekamp and Gao
        n, t = 64, 21
        k, batch = t + 1, 16384 if not args.quick else 2048
        pt = EvalPoint(GF(P), n, False)
        xs_i = [pt(i).value for i in range(n)]
        xs = ntl.pack_vec(xs_i, P)
        msg = synth(batch, k, 60)
        enc = ntl.vandermonde_batch_evaluate_limbs(xs, msg, P)
        rng = np.random.default_rng(61)
        noise = synth(batch, n, 62)
        bad = np.zeros((batch, n), dtype=bool)
        for b in range(batch):
            bad[b, rng.choice(n, t, replace=False)] = True
        words = np.where(bad[:, :, None], noise, enc)
        e_max = (n - t) // 2
        t0 = time.perf_counter()
        coeffs, out_len, status = robust.wb_decode_batch_limbs(xs, words, k, e_max, P)
        wb_s = time.perf_counter() - t0
        assert (status == 0).all() and np.array_equal(coeffs, msg), "cfg3 WB"
        t0 = time.perf_counter()
        coeffs, out_len, status = robust.wb_decode_batch_limbs(xs, words, k, e_max, P)
        wb_s = min(wb_s, time.perf_counter() - t0)
        t0 = time.perf_counter()
        gc, loc, ll, gst = robust.gao_decode_batch_limbs(xs, words, k, P)
        gao_s = time.perf_counter() - t0
        t0 = time.perf_counter()
        gc, loc, ll, gst = robust.gao_decode_batch_limbs(xs, words, k, P)
        gao_s = min(gao_s, time.perf_counter() - t0)
        assert (gst == 0).all() and np.array_equal(gc, msg) and (ll == t + 1).all(), "cfg3 Gao"
        out.append({"config": f"cfg3 n=64 t=21 robust decode, 21 errors/word, batch={batch} (host buffers)",
                    "wb_s": wb_s, "wb_words_per_s": batch / wb_s, "gao_s": gao_s,
                    "gao_words_per_s": batch / gao_s,
                    "wb_GBps_algorithmic": batch * (n + k) * E / wb_s / 1e9})

    # ---- config 1: n=4, t=1, batch_reconstruct of 256 shares, 4 parties in one process
    from sim_net import SimNet

    from honeybadgermpc_b200.batch_reconstruction import batch_reconstruct

    ctx.set_stream(None)
    n, t, count = 4, 1, 256
    fp = GF(P)
    rngp = random.Random(1)
    secrets = [rngp.randrange(P) for _ in range(count)]
    polys = [[s, rngp.randrange(P)] for s in secrets]
    shares = [[(c0 + c1 * (i + 1)) % P for c0, c1 in polys] for i in range(n)]

    async def once():
        net = SimNet(n)
        jobs = [batch_reconstruct([fp(v) for v in shares[i]], P, t, n, i, net.sends[i], net.recvs[i])
                for i in range(n)]
        return await asyncio.gather(*jobs)

    loop = asyncio.new_event_loop()
    res = loop.run_until_complete(once())
    assert all([e.value for e in r] == secrets for r in res)
    t0 = time.perf_counter()
    reps = 5
    for _ in range(reps):
        loop.run_until_complete(once())
    dt = (time.perf_counter() - t0) / reps
    loop.close()
    out.append({"config": "cfg1 n=4 t=1 batch_reconstruct of 256 shares, 4 parties in-process (wall time)",
                "s_per_open_all_parties": dt, "shares_per_s_per_party": count / dt})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
