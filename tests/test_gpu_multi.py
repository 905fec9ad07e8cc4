"""The multi-GPU path (honeybadgermpc_b200.sharding.ShardedReconstructor): one rank per GPU,
batch axis sharded, decoded blocks all-gathered on every rank -- every gather
implementation (fused into the kernel epilogue through the multicast address or peer
stores, side-stream copy kernel, NCCL), multi-part slots and the CUDA-graph form, each
bit-exact against the oracle on every rank.  Needs >= 2 GPUs: the round-end single-GPU
run skips it; ``gpurun --gpus 2|4`` runs it (log committed under profiles/).
``pytest -m gpu``."""

import random
import socket

import pytest
from conftest import BLS12_381_R as P

pytestmark = pytest.mark.gpu


def _rank_main(rank, world, port, results):
    import numpy as np
    import torch
    import torch.distributed as dist

    from honeybadgermpc_b200 import _native
    from honeybadgermpc_b200.ntl import pack_vec, unpack_rows
    from honeybadgermpc_b200.sharding import ShardedReconstructor
    from oracle import hbmpc_oracle as orc

    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world,
                            device_id=torch.device("cuda", rank))
    ok = []
    try:
        dev = torch.device("cuda", rank)
        results[rank] = "started"
        for (n, k, rows, parts) in [(16, 6, 1500, 1), (16, 6, 700, 3), (128, 43, 300, 2)]:
            pt = orc.EvalPoint(P, n, True)
            omega = pack_vec([pt.omega], P)[0]
            zs = sorted(random.Random(n + k).sample(range(n), k))
            rng = np.random.default_rng(1000 * rank + n)
            steps = 2 * parts + parts  # three slots' worth, so slots are reused
            c = rng.integers(0, 2 ** 62, size=(steps, rows, k, 4), dtype=np.uint64)
            xs = [pt(z) for z in zs]
            # the shares: evaluations of the polynomials at the points of zs (oracle arithmetic)
            ys = np.zeros_like(c)
            for s in range(steps):
                ints = unpack_rows(c[s][:64])
                ev = orc.vandermonde_batch_evaluate(xs, ints, P)
                ys[s][:64] = np.frombuffer(b"".join(v.to_bytes(32, "little") for r in ev for v in r),
                                           dtype=np.uint64).reshape(64, k, 4)
            ctx0 = _native.Context(P, device=rank)
            # rows beyond the first 64 of each step: computed with the kernels themselves
            enc = np.zeros((rows, n, 4), np.uint64)
            for s in range(steps):
                ctx0.fft_batch_evaluate(omega, pt.order, c[s], rows, k, n, enc)
                assert np.array_equal(enc[:64][:, zs, :], ys[s][:64])  # kernels agree with the oracle
                ys[s] = enc[:, zs, :]
            y_dev = torch.from_numpy(ys.view(np.int64)).to(dev)
            modes = (["auto", "ce", "mc", "p2p", "bulk", "fused", "fused-barrier", "copy", "nccl"] if k <= 8
                     else ["auto", "mc", "bulk", "copy", "nccl"])
            for gather in modes:
                for graph in (False, True):
                    if graph and gather not in ("auto", "ce", "mc", "p2p", "bulk", "fused"):
                        continue  # CUDA graphs: the device-flag hand-over (no host-issued collective)
                    rec = ShardedReconstructor(P, omega, pt.order, zs, rows, device=rank, depth=2,
                                               gather=gather, parts=parts)
                    ptrs = [y_dev[s].data_ptr() for s in range(steps)]
                    if graph:
                        g = rec.capture(ptrs)
                        with torch.cuda.stream(rec.stream):
                            g.replay()
                    else:
                        for s in range(steps):
                            slot = rec.open(ptrs[s], part=s % parts)
                            if s % parts == parts - 1:
                                rec.finish(slot)
                    rec.drain()
                    torch.cuda.synchronize()
                    dist.barrier()
                    # the last `depth` slot fills: slot contents vs this rank's coefficients
                    fills = steps // parts
                    for f in range(max(0, fills - 2), fills):
                        got = rec.gathered[f % 2].cpu().numpy().view(np.uint64)
                        got = got.reshape(world, parts, rows, k, 4)
                        mine = got[rank]
                        want = c[f * parts:(f + 1) * parts]
                        ok.append(bool(np.array_equal(mine, want)))
                        # every rank holds the same gathered array
                        digest = torch.from_numpy(got.view(np.int64).reshape(world, -1)).sum(dim=1).to(dev)
                        lo, hi = digest.clone(), digest.clone()
                        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
                        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
                        ok.append(bool(torch.equal(lo, hi)))
                    ok.append(rec.mode != "local")
                    del rec
        results[rank] = True if (all(ok) and len(ok) > 0) else f"mismatch: {ok}"
    except Exception:  # noqa: BLE001 - reported to the parent, which asserts
        import traceback

        results[rank] = traceback.format_exc()[-1500:]
        raise
    finally:
        try:
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass


def test_sharded_reconstructor_multi_gpu():
    import torch
    import torch.multiprocessing as mp

    world = min(torch.cuda.device_count(), 4)
    if world < 2:
        pytest.skip("needs at least two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_rank_main, args=(r, world, port, results)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(600)
    assert dict(results) == {r: True for r in range(world)}, dict(results)
    assert all(p.exitcode == 0 for p in procs)


def test_sharded_reconstructor_single_gpu():
    """world = 1: the same object degenerates to a local interpolation into its slot
    (this is what bench.py drives at N = 1), eager and as a CUDA graph"""
    import numpy as np
    import torch

    from honeybadgermpc_b200 import _native
    from honeybadgermpc_b200.ntl import pack_vec, unpack_rows
    from honeybadgermpc_b200.sharding import ShardedReconstructor
    from oracle import hbmpc_oracle as orc

    n, k, rows = 16, 6, 2000
    pt = orc.EvalPoint(P, n, True)
    omega = pack_vec([pt.omega], P)[0]
    zs = [1, 3, 4, 9, 12, 15]
    rng = np.random.default_rng(3)
    c = rng.integers(0, 2 ** 62, size=(4, rows, k, 4), dtype=np.uint64)
    ctx = _native.Context(P, device=0)
    ys = np.zeros_like(c)
    enc = np.zeros((rows, n, 4), np.uint64)
    for s in range(4):
        ctx.fft_batch_evaluate(omega, pt.order, c[s], rows, k, n, enc)
        ys[s] = enc[:, zs, :]
    assert unpack_rows(enc[:8]) == orc.fft_batch_evaluate(unpack_rows(c[3][:8]), pt.omega, P, pt.order, n)
    y_dev = torch.from_numpy(ys.view(np.int64)).cuda()
    for graph in (False, True):
        rec = ShardedReconstructor(P, omega, pt.order, zs, rows, device=0, depth=2)
        assert rec.mode == "local" and rec.world == 1
        ptrs = [y_dev[s].data_ptr() for s in range(4)]
        if graph:
            g = rec.capture(ptrs)
            with torch.cuda.stream(rec.stream):
                g.replay()
        else:
            for p in ptrs:
                rec.open(p)
        rec.drain()
        for s in (2, 3):
            assert np.array_equal(rec.gathered[s % 2].cpu().numpy().view(np.uint64), c[s])
