#!/usr/bin/env python
"""Where does a multi-GPU step go?  Times, per rank: the raw copy-engine peer copy of one decoded
block, the interpolation alone, and open()+finish() sequences (eager and CUDA graph) for each
gather mode.   torchrun --nproc-per-node N tools/gather_probe.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from honeybadgermpc_b200 import _native  # noqa: E402
from honeybadgermpc_b200.field import GF  # noqa: E402
from honeybadgermpc_b200.ntl import pack_vec  # noqa: E402
from honeybadgermpc_b200.polynomial import EvalPoint  # noqa: E402
from honeybadgermpc_b200.sharding import ShardedReconstructor  # noqa: E402

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
dev = torch.device("cuda", local)
n, k, rows, zs = 16, 6, 65536, [1, 3, 4, 9, 12, 15]
pt = EvalPoint(GF(P), n, True)
omega = pack_vec([pt.omega.value], P)[0]
ys = [torch.randint(0, 2 ** 60, (rows, k, 4), dtype=torch.int64, device=dev) for _ in range(6)]
out = {}


def timed(fn, reps, stream):
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    fn(reps)
    e1.record(stream)
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


for gather in ("ce", "mc", "nccl"):
    rec = ShardedReconstructor(P, omega, pt.order, zs, rows, device=local, depth=3, gather=gather)
    if gather == "ce" and rec.handles:
        # raw peer copy of one block on the side stream
        peer = (rank + 1) % world
        dst = int(rec.handles[0].buffer_ptrs[peer]) + rank * rec.block_bytes
        src = rec.own_block_ptr(0)
        import ctypes

        rt = ctypes.CDLL("libcudart.so.12")

        def raw(reps):
            for _ in range(reps):
                rt.cudaMemcpyAsync(ctypes.c_void_p(dst), ctypes.c_void_p(src), ctypes.c_size_t(rec.block_bytes),
                                   3, ctypes.c_void_p(rec.sides[0].cuda_stream))
        raw(3)
        out["raw_ce_copy_us"] = timed(raw, 50, rec.sides[0])

    def interp_only(reps):
        for i in range(reps):
            rec.ctx.fft_batch_interpolate(rec.omega, rec.order, rec.zs, ys[i % 6].data_ptr(), rows,
                                          rec.own_block_ptr(i % 3), _native.MEM_DEVICE)
    interp_only(3)
    out["interp_only_us"] = timed(interp_only, 60, rec.stream)

    def eager(reps):
        for i in range(reps):
            s = rec.open(ys[i % 6].data_ptr(), slot=i % 3)
            rec.finish(s)
        for st in rec.sides:
            rec.stream.wait_stream(st)
        rec.stream.wait_stream(rec.sync)
    eager(6)
    rec.drain()
    out[f"{rec.mode}_eager_us"] = timed(eager, 120, rec.stream)
    rec.drain()
    if gather != "nccl":
        g = rec.capture([ys[i % 6].data_ptr() for i in range(12)])

        def graph(reps):
            with torch.cuda.stream(rec.stream):
                for _ in range(reps // 12):
                    g.replay()
        graph(12)
        out[f"{rec.mode}_graph12_us"] = timed(graph, 240, rec.stream)
    rec.drain()
    del rec
if rank == 0:
    out["world"] = world
    out["block_MB"] = rows * k * 32 / 1e6
    print(json.dumps(out))
dist.barrier()
dist.destroy_process_group()
