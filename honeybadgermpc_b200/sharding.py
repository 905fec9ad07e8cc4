"""Batch-axis sharding of the reconstruction path across the GPUs of one node.

Every row (polynomial) of every primitive is independent (the reference already
parallelises exactly this axis on CPU cores: OpenMP ``prange`` over the batch,
ntl/hbmpc_ntl_helpers.pyx:306-309, :369-374), so each rank processes a
contiguous block of ceil(batch / world) rows with no data-path collective, and
ONE all-gather reassembles the decoded blocks so that every rank holds the full
result (every party learns all opened values).  ``torch.distributed`` is the
plumbing: NCCL over NVLink on GPUs, gloo in the CPU tests.

Two layers:

* ``shard_bounds / all_gather_rows / sharded_apply`` -- backend-neutral helpers
  (``all_gather_into_tensor``; what the gloo tests cover).
* ``ShardedReconstructor`` -- the measured GPU path (``bench.py --gpus N``): the
  gather buffers are symmetric-memory allocations mapped into every process
  (and bound to one NVSwitch multicast address where the fabric offers it); the
  interpolation kernel's epilogue stores each decoded element straight into
  every rank's buffer (``hbg_fft_batch_interpolate_allgather``: one
  ``multimem.st`` per 16 bytes, or ``world`` peer stores), or -- ``gather="copy"``
  -- writes its own block locally and a few-CTA copy kernel pushes it out on a
  side stream.  Slot hand-over is two device-side barriers per step on the side
  stream (everyone may overwrite slot s / everyone's block has landed in slot s),
  so no rank ever stalls on the slowest one inside its compute stream, and the
  whole per-slot sequence can be captured into a CUDA graph (``capture``): the
  host then issues one graph launch per step.
"""

import torch
import torch.distributed as dist

from . import _native


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


class ShardedReconstructor:
    """One rank of a batch-sharded ``fft_batch_interpolate`` whose decoded blocks
    are all-gathered on every rank.

    ``rows``: polynomials per rank and step; ``zs``: the k party indices the
    shares come from; ``omega`` (``uint64[4]``) / ``order``: the evaluation
    domain.  ``depth`` gather slots rotate, so up to ``depth`` steps are in
    flight.  ``gather``:

      ``"auto"``   fused into the kernel epilogue, through the multicast address
                   when there is one, else peer stores
      ``"p2p"``    fused, peer stores only
      ``"copy"``   local store + side-stream copy kernel (``hbg_allgather_block``)
      ``"nccl"``   local store + ``all_gather_into_tensor`` on NCCL's stream

    ``open(y_ptr)`` enqueues one step and returns its slot; ``wait(slot)``
    makes the caller's current stream wait for the gathered result
    ``gathered[slot]`` (``int64[world * rows, k, 4]``); ``release(slot)`` hands
    the slot back."""

    def __init__(self, modulus, omega, order, zs, rows, group=None, device=None, depth=3, gather="auto",
                 copy_ctas=16):
        import numpy as np

        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.rows, self.k, self.order = int(rows), len(zs), int(order)
        self.omega = np.ascontiguousarray(omega, dtype=np.uint64)
        self.zs = np.ascontiguousarray(zs, dtype=np.int32)
        self.copy_ctas = copy_ctas
        self.ctx = _native.Context(modulus, device=self.device.index)
        self.stream = torch.cuda.Stream(device=self.device)
        self.side = torch.cuda.Stream(device=self.device)
        self.ctx.set_stream(self.stream.cuda_stream)
        self.side_ctx = _native.Context(modulus, device=self.device.index)
        self.side_ctx.set_stream(self.side.cuda_stream)
        self.block_bytes = self.rows * self.k * 32
        self.handles, self.gathered = [], []
        self.mode = "nccl" if gather == "nccl" else None
        if self.world > 1 and self.mode is None:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                for _ in range(depth):
                    buf = symm_mem.empty((self.world * self.rows, self.k, 4), dtype=torch.int64,
                                         device=self.device)
                    self.handles.append(symm_mem.rendezvous(buf, self.group))
                    self.gathered.append(buf)
                mc = int(getattr(self.handles[0], "multicast_ptr", 0) or 0)
                if gather == "copy":
                    self.mode = "multimem-copy" if mc else "p2p-copy"
                elif gather == "auto":
                    self.mode = "fused-multimem" if mc else "fused-p2p"
                else:
                    self.mode = "fused-p2p"
            except Exception as exc:  # noqa: BLE001 - no symmetric memory on this build / fabric
                self.fallback_reason = repr(exc)
                self.handles, self.gathered, self.mode = [], [], "nccl"
        if not self.gathered:
            self.mode = "nccl" if self.world > 1 else "local"
            self.gathered = [torch.empty((self.world * self.rows, self.k, 4), dtype=torch.int64,
                                         device=self.device) for _ in range(depth)]
        self.depth = len(self.gathered)
        self.peers = [_native.Context.peer_array(list(h.buffer_ptrs)) for h in self.handles]
        self.mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.handles]
        mk = torch.cuda.Event
        self.ready_ev = [mk() for _ in range(self.depth)]
        self.written_ev = [mk() for _ in range(self.depth)]
        self.done_ev = [mk() for _ in range(self.depth)]
        self.released_ev = [None] * self.depth
        self.pending = [None] * self.depth  # NCCL work handles
        self.step_no = 0

    # -- one step -----------------------------------------------------------------
    def own_block_ptr(self, slot):
        return self.gathered[slot].data_ptr() + self.rank * self.block_bytes

    def open(self, y_ptr, slot=None):
        """enqueue: interpolate ``rows`` polynomials from the device array at ``y_ptr``
        (``[rows, k, 4]`` uint64) and all-gather the result into slot ``slot``"""
        if slot is None:
            slot = self.step_no % self.depth
        self.step_no += 1
        fused = self.mode.startswith("fused")
        if self.handles:
            # (1) everyone has released slot `slot`: its previous contents may be overwritten
            with torch.cuda.stream(self.side):
                if self.released_ev[slot] is not None:
                    self.side.wait_event(self.released_ev[slot])
                self.handles[slot].barrier(channel=0)
                self.ready_ev[slot].record(self.side)
            self.stream.wait_event(self.ready_ev[slot])
        elif self.pending[slot] is not None:
            self.pending[slot].wait()
            self.pending[slot] = None
        if fused:
            self.ctx.fft_batch_interpolate_allgather(
                self.omega, self.order, self.zs, y_ptr, self.rows, self.peers[slot],
                self.mc[slot] if self.mode == "fused-multimem" else 0, self.rank)
        else:
            self.ctx.fft_batch_interpolate(self.omega, self.order, self.zs, y_ptr, self.rows,
                                           self.own_block_ptr(slot), _native.MEM_DEVICE)
        if self.handles:
            self.written_ev[slot].record(self.stream)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.written_ev[slot])
                if not fused:
                    self.side_ctx.allgather_block(
                        self.own_block_ptr(slot), self.block_bytes, self.peers[slot],
                        self.mc[slot] if self.mode == "multimem-copy" else 0,
                        self.rank * self.block_bytes, self.copy_ctas)
                # (2) every rank's block has landed in every buffer
                self.handles[slot].barrier(channel=1)
                self.done_ev[slot].record(self.side)
        elif self.world > 1:
            with torch.cuda.stream(self.stream):
                own = self.gathered[slot][self.rank * self.rows:(self.rank + 1) * self.rows]
                self.pending[slot] = dist.all_gather_into_tensor(self.gathered[slot], own, group=self.group,
                                                                 async_op=True)
        else:
            self.done_ev[slot].record(self.stream)
        return slot

    def wait(self, slot):
        """the caller's current stream waits for slot ``slot`` to be complete"""
        if self.pending[slot] is not None:
            self.pending[slot].wait()
            self.pending[slot] = None
        elif self.handles or self.world == 1:
            torch.cuda.current_stream(self.device).wait_event(self.done_ev[slot])

    def release(self, slot):
        """the caller's current stream is done reading slot ``slot``"""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self.released_ev[slot] = ev

    def drain(self):
        for slot in range(self.depth):
            if self.pending[slot] is not None:
                self.pending[slot].wait()
                self.pending[slot] = None
        self.stream.synchronize()
        self.side.synchronize()

    # -- CUDA graph of a sequence of steps ----------------------------------------
    def capture(self, y_ptrs, begin=None, extra=None, finish=None):
        """Capture ``len(y_ptrs)`` consecutive steps (slot i % depth) into one CUDA
        graph.  ``begin()`` may fork further streams from ``self.stream``, ``extra(i)``
        enqueue more work per step on them, ``finish()`` must join them back.  Returns
        the ``torch.cuda.CUDAGraph``; replaying it costs the host one launch.  NCCL
        mode is not capturable (returns ``None``)."""
        if self.mode == "nccl":
            return None
        self.drain()
        graph = torch.cuda.CUDAGraph()
        self.released_ev = [None] * self.depth
        with torch.cuda.graph(graph, stream=self.stream, capture_error_mode="thread_local"):
            self.side.wait_stream(self.stream)  # fork: the side stream is part of the capture
            if begin is not None:
                begin()
            for i, y in enumerate(y_ptrs):
                if extra is not None:
                    extra(i)
                self.open(y, slot=i % self.depth)
            if finish is not None:
                finish()
            self.stream.wait_stream(self.side)
        return graph
