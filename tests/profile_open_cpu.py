"""Where does a large open spend its host time?  batch_reconstruct at n = 16, t = 5 with all 16
parties in one process, the kernels replaced by the C++ CPU restatement (oracle/cpu_ref.cpp)
behind the native Context interface, so that what remains is the Python around the C-ABI.
Not a test (no test_ prefix); run by hand:   python tests/profile_open_cpu.py [B] [profile]"""
import asyncio
import cProfile
import os
import pstats
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, HERE)

import numpy as np  # noqa: E402
from sim_net import SimNet  # noqa: E402

from honeybadgermpc_b200 import _native, ntl  # noqa: E402
from honeybadgermpc_b200.batch_reconstruction import batch_reconstruct  # noqa: E402
from honeybadgermpc_b200.field import GF  # noqa: E402
from oracle import cpu_ref  # noqa: E402

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


class CpuRefContext:
    def __init__(self, p):
        self.p, self.ref = int(p), cpu_ref.CpuRef()

    @staticmethod
    def _pts(xs):  # uint64[n, 4] -> ints (cpu_ref takes the points as Python ints)
        return ntl.unpack_rows(np.ascontiguousarray(xs)[None])[0]

    def vandermonde_batch_evaluate(self, xs, polys, batch, d, out, mem=0):
        out[...] = self.ref.vandermonde_batch_evaluate_limbs(self._pts(xs), polys, self.p)

    def vandermonde_batch_interpolate(self, xs, ys, batch, out, mem=0):
        out[...] = self.ref.vandermonde_batch_interpolate_limbs(self._pts(xs), ys, self.p)


_ctx = CpuRefContext(P)
_native.get_context = lambda modulus, device=None: _ctx

n, t = 16, 5
B = int(sys.argv[1]) if len(sys.argv) > 1 else 6 * 8192
fp = GF(P)
rng = np.random.default_rng(1)
polys = rng.integers(0, 2 ** 62, size=(B, t + 1, 4), dtype=np.uint64)
secrets = ntl.unpack_rows(np.ascontiguousarray(polys[:, :1, :]).reshape(1, B, 4))[0]
xs = ntl.pack_vec(list(range(1, n + 1)), P)
shares_l = ntl.vandermonde_batch_evaluate_limbs(xs, polys, P)
shares = [ntl.wrap_elements(np.ascontiguousarray(shares_l[:, i, :]), fp) for i in range(n)]


async def once(wire):
    net = SimNet(n)
    jobs = [batch_reconstruct(shares[i], P, t, n, i, net.sends[i], net.recvs[i], wire=wire) for i in range(n)]
    return await asyncio.gather(*jobs)


for wire in ("ints", "limbs"):
    loop = asyncio.new_event_loop()
    res = loop.run_until_complete(once(wire))
    assert [e.value for e in res[0]] == secrets
    t0 = time.perf_counter()
    loop.run_until_complete(once(wire))
    dt = time.perf_counter() - t0
    print(f"wire={wire}: B={B} shares, n={n}: {dt / n * 1e3:.1f} ms per party (CPU kernels included)")
    if len(sys.argv) > 2:
        pr = cProfile.Profile()
        pr.enable()
        loop.run_until_complete(once(wire))
        pr.disable()
        pstats.Stats(pr).sort_stats("tottime").print_stats(14)
    loop.close()
