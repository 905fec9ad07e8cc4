#!/usr/bin/env python
"""Duration of one tensor-core launch (cfg2 encode / interpolate, 65 536 rows unless --batch) as a
function of the CTAs it may use (hbg_ctx_set_sm_limit): separates the fixed cost of a launch
from the cost per 128-row tile.  One JSON line per (op, limit)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch  # noqa: E402
import bench  # noqa: E402
from honeybadgermpc_b200 import _native  # noqa: E402
from honeybadgermpc_b200.field import GF  # noqa: E402
from honeybadgermpc_b200.polynomial import EvalPoint  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--limits", default="148,128,111,104,74,52,44,37")
    ap.add_argument("--reps", type=int, default=60)
    a = ap.parse_args()
    P, K, N = bench.P, bench.K, bench.N_PARTIES
    pt = EvalPoint(GF(P), N, True)
    from honeybadgermpc_b200.ntl import pack_vec
    omega = pack_vec([pt.omega.value], P)[0]
    ctx = _native.Context(P, device=0)
    st = torch.cuda.Stream()
    ctx.set_stream(st.cuda_stream)
    sets = max(2, int(np.ceil(300e6 / (a.batch * 32 * (K + N)))))
    cs = [torch.from_numpy(bench.synth(a.batch, K, 7 + s).view(np.int64)).cuda() for s in range(sets)]
    es = [torch.empty((a.batch, N, 4), dtype=torch.int64, device="cuda") for _ in range(sets)]
    zs = np.asarray(bench.ZS, dtype=np.int32)
    zt = torch.tensor(bench.ZS, device="cuda")
    with torch.cuda.stream(st):
        for s in range(sets):
            ctx.fft_batch_evaluate(omega, pt.order, cs[s].data_ptr(), a.batch, K, N, es[s].data_ptr(), _native.MEM_DEVICE)
        ys = [e.index_select(1, zt).contiguous() for e in es]
        os_ = [torch.empty_like(y) for y in ys]
    st.synchronize()

    def enc(s):
        ctx.fft_batch_evaluate(omega, pt.order, cs[s].data_ptr(), a.batch, K, N, es[s].data_ptr(), _native.MEM_DEVICE)

    def dec(s):
        ctx.fft_batch_interpolate(omega, pt.order, zs, ys[s].data_ptr(), a.batch, os_[s].data_ptr(), _native.MEM_DEVICE)

    for name, fn in (("encode", enc), ("interpolate", dec)):
        for lim in [int(x) for x in a.limits.split(",")]:
            ctx.set_sm_limit(lim)
            for s in range(sets):
                fn(s)
            st.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(a.reps)]
            with torch.cuda.stream(st):
                for i in range(a.reps):
                    ev[i][0].record(st)
                    fn(i % sets)
                    ev[i][1].record(st)
            st.synchronize()
            t = sorted(x.elapsed_time(y) * 1e3 for x, y in ev)
            tiles = (a.batch + 127) // 128
            print(json.dumps({"op": name, "sm_limit": lim, "rounds": -(-tiles // lim), "us_median": round(t[len(t) // 2], 2),
                              "us_min": round(t[0], 2), "kernel": ctx.last_kernel()}), flush=True)
    assert torch.equal(os_[0], cs[0])


if __name__ == "__main__":
    main()
