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


def gather_mode(gather, world, multicast):
    """The transport behind ``gather=`` for ``world`` ranks (``multicast``: the symmetric-memory
    allocation has an NVSwitch multicast address).  ``"auto"`` follows the measurements
    (profiles/r2_gather_modes_n4.txt, r2g_*): 2 ranks -- copy engines (30 us per cfg2 step; copy
    kernels 38-57); 3-4 ranks -- the unicast TMA bulk-copy kernel on 16 SMs the compute kernels leave
    free (70 us at 4 ranks; multicast kernel 83, copy engines 95: unicast receives world-1 blocks, a
    multicast store world -- the sender's own block comes back through the switch); 5 ranks and more
    -- the multicast kernel (154 us at 8, copy engines 238: one block leaves the GPU instead of
    seven), or the bulk-copy kernel where there is no multicast address."""
    if gather == "auto":
        if world <= 2:
            return "ce-copy-signal"
        if world <= 4 or not multicast:
            return "bulk-copy-signal"
        return "multimem-copy-signal"
    return {"ce": "ce-copy-signal",
            "mc": "multimem-copy-signal" if multicast else "p2p-copy-signal",
            "p2p": "p2p-copy-signal",
            "bulk": "bulk-copy-signal",
            "fused": "fused-multimem-signal" if multicast else "fused-p2p-signal",
            "fused-barrier": "fused-multimem" if multicast else "fused-p2p",
            "copy": "multimem-copy" if multicast else "p2p-copy"}[gather]


class ShardedReconstructor:
    """One rank of a batch-sharded ``fft_batch_interpolate`` whose decoded blocks
    are all-gathered on every rank.

    ``rows``: polynomials per rank and step; ``zs``: the k party indices the
    shares come from; ``omega`` (``uint64[4]``) / ``order``: the evaluation
    domain.  ``depth`` gather slots rotate, so up to ``depth`` steps are in
    flight.  ``gather``:

      ``"auto"``     ``"ce"`` for two ranks, ``"mc"`` from four ranks on (measured).
      ``"ce"``       the interpolation kernel writes its block into its own
                     slot of the local gather buffer; on a side stream the COPY ENGINES
                     push it to every peer (``hbg_allgather_block_ce``: one
                     ``cudaMemcpyAsync`` per peer into the symmetric-memory mapping, no
                     SM touches the payload) and the slot hand-over runs on DEVICE flags
                     in symmetric memory (``hbg_gather_wait`` / ``hbg_gather_release``):
                     no host-issued barrier, nothing holds compute SMs while NVLink drains
      ``"mc"``       the same with a copy KERNEL storing through the NVSwitch multicast
                     address (``multimem.st``: the block leaves this GPU once and the
                     switch replicates it; ``hbg_allgather_block_signal``)
      ``"p2p"``      the copy kernel with peer stores only
      ``"bulk"``     a copy kernel that moves the block with TMA bulk copies (global -> shared ->
                     every peer; ``hbg_allgather_block_bulk``): unicast like ``"ce"``, one launch
                     like ``"mc"``, a few SMs
      ``"fused"``    the kernel epilogue stores into every rank's buffer itself
                     (``hbg_fft_batch_interpolate_allgather``: full 128-byte lines through
                     the multicast address or to every peer, no second pass over the block),
                     hand-over by the same device flags (``hbg_gather_fence``)
      ``"fused-barrier"``  the same with two symmetric-memory barriers per step (round 1)
      ``"copy"``     copy kernel + the two barriers (round 1's protocol)
      ``"nccl"``     local store + ``all_gather_into_tensor`` on NCCL's stream

    ``open(y_ptr)`` enqueues one step and returns its slot; ``wait(slot)``
    makes the caller's current stream wait for the gathered result
    ``gathered[slot]`` (``int64[world * parts * rows, k, 4]``); ``release(slot)``
    hands the slot back (``finish(slot)`` = both, on an internal stream: what a
    producer-only benchmark calls).

    ``parts`` > 1: a slot is filled by ``parts`` consecutive steps of ``rows``
    polynomials each (step ``part`` of rank r lands at row ``(r * parts + part) * rows``),
    so the gather of one piece overlaps the interpolation of the next -- the shape of a
    strong-scaling job where the whole batch is one result (BASELINE configs[4])."""

    def __init__(self, modulus, omega, order, zs, rows, group=None, device=None, depth=3, gather="auto",
                 copy_ctas=64, parts=1, compute_sms=0):
        import numpy as np

        self.group = group if group is not None else (dist.group.WORLD if dist.is_initialized() else None)
        self.world = dist.get_world_size(self.group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(self.group) if dist.is_initialized() else 0
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.rows, self.k, self.order, self.parts = int(rows), len(zs), int(order), int(parts)
        self.omega = np.ascontiguousarray(omega, dtype=np.uint64)
        self.zs = np.ascontiguousarray(zs, dtype=np.int32)
        self.copy_ctas = copy_ctas
        mk_stream = lambda: torch.cuda.Stream(device=self.device)  # noqa: E731
        mk_ctx = lambda: _native.Context(modulus, device=self.device.index)  # noqa: E731
        self.stream, self.sync = mk_stream(), mk_stream()
        # one side stream per slot: the chain [wait for the release of the slot -> copies ->
        # publish] of a step costs two engine switches (~8 us each, measured) on top of the
        # copy itself; on separate streams the chains of consecutive steps overlap and only
        # the copies themselves queue up on the link
        self.sides = [mk_stream() for _ in range(depth)]
        self.ctx, self.sync_ctx = mk_ctx(), mk_ctx()
        self.side_ctxs = [mk_ctx() for _ in range(depth)]
        self.user_ctx = mk_ctx()  # wait()/release() on caller streams
        # compute_sms > 0: the interpolation uses at most that many SMs (hbg_ctx_set_sm_limit), so
        # the copy kernel of the gather finds free SMs instead of queueing behind a full-GPU launch
        self.compute_sms = int(compute_sms)
        if self.compute_sms > 0:
            self.ctx.set_sm_limit(self.compute_sms)
        self.ctx.set_stream(self.stream.cuda_stream)
        for c, st in zip(self.side_ctxs, self.sides):
            c.set_stream(st.cuda_stream)
        self.sync_ctx.set_stream(self.sync.cuda_stream)
        self.block_bytes = self.rows * self.k * 32
        self.handles, self.gathered, self.flag_handle, self.flags = [], [], None, None
        self.mode = "nccl" if gather == "nccl" else None
        shape = (self.world * self.parts * self.rows, self.k, 4)
        if self.world > 1 and self.mode is None:
            try:
                import torch.distributed._symmetric_memory as symm_mem

                for _ in range(depth):
                    buf = symm_mem.empty(shape, dtype=torch.int64, device=self.device)
                    self.handles.append(symm_mem.rendezvous(buf, self.group))
                    self.gathered.append(buf)
                self.flags = symm_mem.empty((max(64, depth * (2 * self.world + 2)),), dtype=torch.int32,
                                            device=self.device)
                self.flags.zero_()
                self.flag_handle = symm_mem.rendezvous(self.flags, self.group)
                torch.cuda.synchronize(self.device)
                dist.barrier(self.group)  # every rank's flags are zero before anyone signals
                mc = int(getattr(self.handles[0], "multicast_ptr", 0) or 0)
                self.mode = gather_mode(gather, self.world, bool(mc))
            except Exception as exc:  # noqa: BLE001 - no symmetric memory on this build / fabric
                self.fallback_reason = repr(exc)
                self.handles, self.gathered, self.mode = [], [], "nccl"
        if not self.gathered:
            self.mode = "nccl" if self.world > 1 else "local"
            self.gathered = [torch.empty(shape, dtype=torch.int64, device=self.device) for _ in range(depth)]
        self.depth = len(self.gathered)
        self.signal = self.mode.endswith("signal")
        if self.mode == "bulk-copy-signal" and gather == "auto":
            # two working threads per CTA: 16 CTAs saturate the links; the compute kernels leave
            # them their SMs (persistent tensor-core CTAs own a whole SM each)
            self.copy_ctas = min(self.copy_ctas, 16)
            if self.compute_sms == 0:
                sms = torch.cuda.get_device_properties(self.device).multi_processor_count
                self.compute_sms = sms - self.copy_ctas
                self.ctx.set_sm_limit(self.compute_sms)
        self.peers = [_native.Context.peer_array(list(h.buffer_ptrs)) for h in self.handles]
        self.mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in self.handles]
        self.flag_peers = _native.Context.peer_array(list(self.flag_handle.buffer_ptrs)) \
            if self.flag_handle is not None else None
        mk = torch.cuda.Event
        self.ready_ev = [mk() for _ in range(self.depth)]
        self.written_ev = [mk() for _ in range(self.depth)]
        self.done_ev = [mk() for _ in range(self.depth)]
        self.released_ev = [None] * self.depth
        self.pending = [None] * self.depth  # NCCL work handles
        self.step_no = 0

    # -- one step -----------------------------------------------------------------
    def own_block_ptr(self, slot, part=0):
        return self.gathered[slot].data_ptr() + (self.rank * self.parts + part) * self.block_bytes

    def open(self, y_ptr, slot=None, part=0):
        """enqueue: interpolate ``rows`` polynomials from the device array at ``y_ptr``
        (``[rows, k, 4]`` uint64) and all-gather the result into slot ``slot``"""
        if slot is None:
            slot = (self.step_no // self.parts) % self.depth
        self.step_no += 1
        fused = self.mode.startswith("fused")
        first, last = part == 0, part == self.parts - 1
        block_row = self.rank * self.parts + part  # in units of `rows`
        if first and self.released_ev[slot] is not None:  # the local reader of the previous fill
            self.stream.wait_event(self.released_ev[slot])
        side, side_ctx = self.sides[slot % len(self.sides)], self.side_ctxs[slot % len(self.sides)]
        if self.handles and first and not self.signal:
            # (1) everyone has released slot `slot`: its previous contents may be overwritten
            with torch.cuda.stream(side):
                self.handles[slot].barrier(channel=0)
                self.ready_ev[slot].record(side)
            self.stream.wait_event(self.ready_ev[slot])
        elif not self.handles and first and self.pending[slot] is not None:
            self.pending[slot].wait()
            self.pending[slot] = None
        if fused:
            # rank-relative placement: the kernel stores row r at gather_row0 + r with
            # gather_row0 = rank * rows, so shift the peer bases by the part's offset
            base = (block_row - self.rank) * self.block_bytes
            peers = self.peers[slot] if base == 0 else _native.Context.peer_array(
                [int(p) + base for p in self.handles[slot].buffer_ptrs])
            mc = self.mc[slot] + base if self.mode.startswith("fused-multimem") else 0
            if self.signal and first:  # every rank has released the slot
                self.ctx.gather_fence(self.flag_peers, self.rank, self.depth, slot, self.parts, 0)
            self.ctx.fft_batch_interpolate_allgather(self.omega, self.order, self.zs, y_ptr, self.rows,
                                                     peers, mc, self.rank)
            if self.signal:            # this rank's part has landed in every buffer
                self.ctx.gather_fence(self.flag_peers, self.rank, self.depth, slot, self.parts, 1)
        else:
            self.ctx.fft_batch_interpolate(self.omega, self.order, self.zs, y_ptr, self.rows,
                                           self.own_block_ptr(slot, part), _native.MEM_DEVICE)
        if self.handles:
            self.written_ev[slot].record(self.stream)
            side.wait_event(self.written_ev[slot])
            use_mc = self.mc[slot] if self.mode.startswith("multimem") else 0
            if fused and self.signal:
                pass  # the kernel stored into every buffer and the fences above did the hand-over
            elif self.mode == "bulk-copy-signal":
                side_ctx.allgather_block_bulk(
                    self.own_block_ptr(slot, part), self.block_bytes, self.peers[slot],
                    block_row * self.block_bytes, self.rank, self.copy_ctas, self.flag_peers, self.depth, slot,
                    self.parts, first)
            elif self.mode == "ce-copy-signal":
                side_ctx.allgather_block_ce(
                    self.own_block_ptr(slot, part), self.block_bytes, self.peers[slot],
                    block_row * self.block_bytes, self.rank, self.flag_peers, self.depth, slot, self.parts,
                    first)
            elif self.signal:
                side_ctx.allgather_block_signal(
                    self.own_block_ptr(slot, part), self.block_bytes, self.peers[slot], use_mc,
                    block_row * self.block_bytes, self.rank, self.copy_ctas, self.flag_peers, self.depth,
                    slot, self.parts, first)
            else:
                with torch.cuda.stream(side):
                    if not fused:
                        side_ctx.allgather_block(
                            self.own_block_ptr(slot, part), self.block_bytes, self.peers[slot], use_mc,
                            block_row * self.block_bytes, self.copy_ctas)
                    if last:
                        # (2) every rank's blocks have landed in every buffer
                        self.handles[slot].barrier(channel=1)
                        self.done_ev[slot].record(side)
        elif self.world > 1:
            if last:
                with torch.cuda.stream(self.stream):
                    per = self.parts * self.rows
                    own = self.gathered[slot][self.rank * per:(self.rank + 1) * per]
                    self.pending[slot] = dist.all_gather_into_tensor(self.gathered[slot], own,
                                                                     group=self.group, async_op=True)
        elif last:
            self.done_ev[slot].record(self.stream)
        return slot

    def _wait_on(self, ctx, stream, slot):
        stream.wait_event(self.written_ev[slot])  # this rank's own block
        ctx.gather_wait(self.flag_peers, self.rank, self.depth, slot, self.parts)

    def wait(self, slot):
        """the caller's current stream waits for slot ``slot`` to be complete"""
        cur = torch.cuda.current_stream(self.device)
        if self.pending[slot] is not None:
            self.pending[slot].wait()
            self.pending[slot] = None
        elif self.signal:
            self.user_ctx.set_stream(cur.cuda_stream)
            self._wait_on(self.user_ctx, cur, slot)
        elif self.handles or self.world == 1:
            cur.wait_event(self.done_ev[slot])

    def release(self, slot):
        """the caller's current stream is done reading slot ``slot``"""
        cur = torch.cuda.current_stream(self.device)
        if self.signal:
            self.user_ctx.set_stream(cur.cuda_stream)
            self.user_ctx.gather_release(self.flag_peers, self.rank, self.depth, slot)
        ev = torch.cuda.Event()
        ev.record(cur)
        self.released_ev[slot] = ev

    def finish(self, slot):
        """wait + release on an internal stream (a producer with no reader of its own: the
        benchmark).  In the barrier modes the hand-over is already part of ``open``."""
        if not self.signal:
            return
        self._wait_on(self.sync_ctx, self.sync, slot)
        self.sync_ctx.gather_release(self.flag_peers, self.rank, self.depth, slot)
        self.done_ev[slot].record(self.sync)
        self.released_ev[slot] = self.done_ev[slot]  # the next fill's kernel must not start earlier

    def drain(self):
        for slot in range(self.depth):
            if self.pending[slot] is not None:
                self.pending[slot].wait()
                self.pending[slot] = None
        self.stream.synchronize()
        for st in self.sides:
            st.synchronize()
        self.sync.synchronize()

    # -- CUDA graph of a sequence of steps ----------------------------------------
    def capture(self, y_ptrs, begin=None, extra=None, finish=None, auto_finish=True):
        """Capture ``len(y_ptrs)`` consecutive steps (slot (i // parts) % depth) into one CUDA
        graph.  ``begin()`` may fork further streams from ``self.stream``, ``extra(i)``
        enqueue more work per step on them, ``finish()`` must join them back.  With
        ``auto_finish`` every filled slot is waited for and released inside the graph.
        Returns the ``torch.cuda.CUDAGraph``; replaying it costs the host one launch (the
        hand-over flags are counted on the device, so replays need no host bookkeeping).
        NCCL mode is not capturable (returns ``None``)."""
        if self.mode == "nccl":
            return None
        # warm-up outside the capture: the first call for a point set builds its constants
        # (cudaMalloc + copy) and opts the kernel in to large shared memory, none of which may
        # happen while a stream is capturing.  Purely local: no flag is touched.
        self.ctx.fft_batch_interpolate(self.omega, self.order, self.zs, y_ptrs[0], self.rows,
                                       self.own_block_ptr(0, 0), _native.MEM_DEVICE)
        self.drain()
        graph = torch.cuda.CUDAGraph()
        self.released_ev = [None] * self.depth
        with torch.cuda.graph(graph, stream=self.stream, capture_error_mode="thread_local"):
            for st in self.sides:  # fork: the side streams are part of the capture
                st.wait_stream(self.stream)
            self.sync.wait_stream(self.stream)
            if begin is not None:
                begin()
            for i, y in enumerate(y_ptrs):
                if extra is not None:
                    extra(i)
                slot, part = (i // self.parts) % self.depth, i % self.parts
                self.open(y, slot=slot, part=part)
                if auto_finish and part == self.parts - 1:
                    self.finish(slot)
            if finish is not None:
                finish()
            for st in self.sides:
                self.stream.wait_stream(st)
            self.stream.wait_stream(self.sync)
        self.released_ev = [None] * self.depth  # events recorded inside a capture are not usable outside
        return graph
