"""wall time of batch_reconstruct at n=16, t=5 with B opened shares per call (16 parties in one process)"""
import asyncio, os, random, sys, time, cProfile, pstats
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from sim_net import SimNet
from honeybadgermpc_b200.batch_reconstruction import batch_reconstruct
from honeybadgermpc_b200.field import GF
from honeybadgermpc_b200 import ntl
P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
n, t = 16, 5
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6 * 8192
fp = GF(P)
rng = np.random.default_rng(1)
secrets = ntl.unpack_rows(rng.integers(0, 2**62, size=(1, B, 4), dtype=np.uint64))[0]
coef = [ntl.unpack_rows(rng.integers(0, 2**62, size=(1, B, 4), dtype=np.uint64))[0] for _ in range(t)]
xs = ntl.pack_vec(list(range(1, n + 1)), P)
polys = ntl.pack_rows([[secrets[b]] + [coef[j][b] for j in range(t)] for b in range(B)], t + 1, P)
shares_l = ntl.vandermonde_batch_evaluate_limbs(xs, polys, P)        # [B][n][4]
shares = [[fp(v) for v in ntl.unpack_rows(np.ascontiguousarray(shares_l[:, i:i+1, :]).reshape(1, B, 4))[0]] for i in range(n)]

async def once(wire):
    net = SimNet(n)
    jobs = [batch_reconstruct(shares[i], P, t, n, i, net.sends[i], net.recvs[i], wire=wire) for i in range(n)]
    return await asyncio.gather(*jobs)

import json
from honeybadgermpc_b200 import reed_solomon as rs

out_lines = []
for mode in ("device", "host"):
    # "host": round 1's IncrementalDecoder (columns as host limb arrays, every decode / encode a
    # HBG_MEM_HOST round trip, numpy compare); "device": columns resident in HBM, fused
    # decode+re-encode kernel, compare kernel, flags + final rows come back
    rs.DEVICE_MIN_BATCH = 256 if mode == "device" else 1 << 60
    for wire in ("ints", "limbs"):
        loop = asyncio.new_event_loop()
        res = loop.run_until_complete(once(wire))
        assert [e.value for e in res[0]] == secrets
        for key in rs._DeviceColumns.totals:
            rs._DeviceColumns.totals[key] = 0
        reps = 3
        t0 = time.perf_counter()
        for _ in range(reps):
            loop.run_until_complete(once(wire))
        dt = (time.perf_counter() - t0) / reps
        tot = dict(rs._DeviceColumns.totals)
        per = {k: v / max(1, tot["decoders"]) for k, v in tot.items() if k != "decoders"}
        line = {"tool": "bench_protocol", "incremental_decoder": mode, "wire": wire, "n": n, "t": t,
                "shares_per_open": B, "ms_per_party_per_open": dt / n * 1e3,
                "shares_per_s_per_party": B / (dt / n),
                "per_decoder": per if mode == "device" else None,
                "decoders_per_open_per_party": 2}
        out_lines.append(line)
        print(json.dumps(line), flush=True)
        if wire == "limbs" and len(sys.argv) > 2:
            pr = cProfile.Profile(); pr.enable(); loop.run_until_complete(once(wire)); pr.disable()
            pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
        loop.close()
