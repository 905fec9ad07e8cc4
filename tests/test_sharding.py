"""The N > 1 host logic (batch sharding + one all-gather) on CPU: world_size 2,
gloo backend, two spawned processes."""

import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
from conftest import ROOT

sys.path.insert(0, ROOT)
from honeybadgermpc_b200.sharding import all_gather_rows, gather_mode, shard_bounds, sharded_apply  # noqa: E402


def test_shard_bounds_cover_the_batch():
    for batch in (0, 1, 5, 16, 17, 1000):
        for world in (1, 2, 3, 8):
            blocks = [shard_bounds(batch, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == batch
            for (a, b), (c, d) in zip(blocks, blocks[1:]):
                assert b == c and a <= b
            per = -(-batch // world) if batch else 0
            assert all(hi - lo <= per for lo, hi in blocks)


def test_gather_mode_policy():
    """which transport `gather="auto"` picks per world size, with and without a multicast address"""
    assert gather_mode("auto", 2, True) == gather_mode("auto", 2, False) == "ce-copy-signal"
    assert gather_mode("auto", 3, True) == gather_mode("auto", 4, False) == "bulk-copy-signal"
    assert gather_mode("auto", 8, True) == "multimem-copy-signal"
    assert gather_mode("auto", 8, False) == "bulk-copy-signal"
    for name in ("ce", "mc", "p2p", "bulk", "fused", "fused-barrier", "copy"):
        for mc in (True, False):
            mode = gather_mode(name, 4, mc)
            assert ("multimem" in mode) <= mc, (name, mc, mode)  # multimem modes need the address
            assert mode.endswith("signal") == (name not in ("fused-barrier", "copy"))
    with pytest.raises(KeyError):
        gather_mode("bogus", 2, True)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, batch, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(7)
        rows = torch.randint(0, 2 ** 62, (batch, 6, 4), generator=g, dtype=torch.int64)

        def fake_decode(block):  # a row-wise map, like decode_batch_limbs
            return block.flip(1) + 1

        full = sharded_apply(fake_decode, rows)
        assert torch.equal(full, fake_decode(rows))
        lo, hi = shard_bounds(batch, world, rank)
        again = all_gather_rows(rows[lo:hi], batch)
        assert torch.equal(again, rows)
        results[rank] = True
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("batch", [10, 11, 1])
def test_two_ranks_gloo(batch):
    world = 2
    port = _free_port()
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    procs = [ctx.Process(target=_worker, args=(r, world, port, batch, results)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert dict(results) == {0: True, 1: True}
