"""Batch-axis sharding of the reconstruction path across the GPUs of one node.

Every row (polynomial) of every primitive is independent (the reference already
parallelises exactly this axis on CPU cores: OpenMP ``prange`` over the batch,
ntl/hbmpc_ntl_helpers.pyx:306-309, :369-374), so each rank processes a
contiguous block of ceil(batch / world) rows with no data-path collective, and
ONE all-gather reassembles the decoded blocks so that every rank holds the full
result (every party learns all opened values).  ``torch.distributed`` is the
plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.
"""

import torch
import torch.distributed as dist


def shard_bounds(batch, world, rank):
    """Contiguous block [lo, hi) of rank ``rank``; blocks differ by at most the
    padding of the last one."""
    per = (batch + world - 1) // world if world > 0 else batch
    lo = min(batch, rank * per)
    hi = min(batch, lo + per)
    return lo, hi


def all_gather_rows(local, batch, group=None, out=None, async_op=False):
    """``local``: this rank's block ``[rows_local, ...]`` (torch tensor, int64
    limbs; CUDA for NCCL, CPU for gloo).  Returns the full ``[batch, ...]``
    tensor on every rank.  Blocks are padded to ceil(batch/world) rows so one
    ``all_gather_into_tensor`` moves everything.

    ``out``: optional preallocated ``[world * ceil(batch/world), ...]`` buffer.
    ``async_op=True`` returns ``(tensor, work)``: the collective runs on the
    communication stream while the caller's stream keeps decoding the next
    block; call ``work.wait()`` before reading ``tensor`` or reusing ``local``."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return (local, None) if async_op else local
    per = (batch + world - 1) // world
    if local.shape[0] < per:
        pad = torch.zeros((per - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
        local = torch.cat([local, pad], dim=0)
    if out is None:
        out = torch.empty((world * per,) + tuple(local.shape[1:]), dtype=local.dtype,
                          device=local.device)
    work = dist.all_gather_into_tensor(out, local.contiguous(), group=group, async_op=async_op)
    return (out[:batch], work) if async_op else out[:batch]


def sharded_apply(fn, rows, group=None):
    """Apply a row-wise batch function (e.g. a decoder's ``decode_batch_limbs``)
    to this rank's block of ``rows`` and gather the results."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    lo, hi = shard_bounds(rows.shape[0], world, rank)
    return all_gather_rows(fn(rows[lo:hi]), rows.shape[0], group)
