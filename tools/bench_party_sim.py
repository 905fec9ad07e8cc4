#!/usr/bin/env python
"""Opening throughput of the n-parties-on-n-GPUs simulation (party_sim.py): every rank is a
party, R1 = all-to-all, R2 = all-gather over NCCL.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \\
        --master-port 29533 tools/bench_party_sim.py [--batch 65536] [--steps 50]

n = N parties, t = (N - 1) // 3.  Prints one JSON line on rank 0."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from honeybadgermpc_b200 import party_sim  # noqa: E402

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--no-check", action="store_true", help="skip the re-encode + compare of both rounds")
    ap.add_argument("--byzantine", type=int, default=0,
                    help="the last K ranks (K <= t) send noise in both rounds: every opening takes the "
                         "robust-decoder fallback")
    args = ap.parse_args()
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, t = world, (world - 1) // 3
    codec = party_sim.CudaCodec(P, n)
    # shares of `batch` secrets: random degree-t polynomials evaluated at x_rank, built on the device
    # from common coefficients (same seed on every rank)
    rng = np.random.default_rng(1234)
    coeffs = rng.integers(0, 2 ** 62, size=(args.batch, t + 1, 4), dtype=np.int64)
    coeffs[:, :, 3] >>= 2
    dc = torch.from_numpy(coeffs).cuda()
    shares = codec.encode(dc)[:, rank].contiguous()        # f_b(x_rank)
    secrets = dc[:, 0].contiguous()
    assert args.byzantine <= t, "at most t faulty parties"
    bad = rank >= world - args.byzantine
    info = {}
    kw = dict(check=not args.no_check, byzantine=bad, info=info)
    for _ in range(args.warmup):
        got, ok = party_sim.batch_reconstruct_collective(shares, t, codec, **kw)
    if not bad:
        assert ok and torch.equal(got, secrets), "opening failed"
        assert info["errors"] == list(range(world - args.byzantine, world)), info
    torch.cuda.synchronize()
    dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        got, ok = party_sim.batch_reconstruct_collective(shares, t, codec, **kw)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device="cuda", dtype=torch.float64)
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        print(json.dumps({"tool": "bench_party_sim", "n_parties": n, "t": t, "shares_per_open": args.batch,
                          "ms_per_open": float(ms.item()),
                          "shares_opened_per_s": args.batch / (float(ms.item()) * 1e-3),
                          "check": not args.no_check, "byzantine_parties": args.byzantine,
                          "robust_rounds_per_open": info.get("robust_rounds"), "errors_found": info.get("errors"),
                          "note": "every party (GPU) learns all opened values; R1 all-to-all + R2 all-gather"}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
