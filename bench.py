#!/usr/bin/env python
"""Headline benchmark: GF(p) share reconstructions/sec at n=16, t=5
(BASELINE.json configs[1]: NTT encode + interpolate, batch = 65 536 polynomials
of t+1 = 6 shares each, BLS12-381 scalar field, synthetic random shares).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]

One "step" = one pass of the hot path over one batch:
    encode      c[batch][6]  --NTT-16-->  e[batch][16]     (fft_batch_evaluate)
    interpolate y[batch][6]  ---------->  r[batch][6]      (fft_batch_interpolate)
    (N > 1)     all-gather of r over NCCL
`value` counts opened shares: batch * (t+1) per step per GPU (every recovered
coefficient is one reconstructed secret, batch_reconstruction.py:117,158).

Prints ONE JSON line (see the task contract): value, roofline (dominant kernel,
CUDA-event timed), e2e (C-ABI with pinned HOST buffers, copies inside the timed
region), cpu_baseline (oracle/cpu_ref.cpp on this box's cores), clocks.
"""

import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

P = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
N_PARTIES, T = 16, 5
K = T + 1
ZS = [1, 3, 4, 9, 12, 15]  # the scattered z set of SURVEY.md section 8(d)
METRIC = "GF(p) share reconstructions/sec at n=16,t=5"
UNIT = "shares/s"
E = 32  # bytes per field element


def synth(batch, width, seed):
    """uniform in [0, 2^254) subset of [0, p): canonical residues as uint64[batch,width,4]"""
    import numpy as np

    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2 ** 63, size=(batch, width, 4), dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(batch, width, 4), dtype=np.uint64)
    a[:, :, 3] >>= np.uint64(2)
    return a


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons of one GPU while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self.stop_flag = threading.Event()
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:  # noqa: BLE001
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
            "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
            "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
            "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
        }
        while not self.stop_flag.is_set():
            try:
                self.samples.append((time.perf_counter(), nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:  # noqa: BLE001
                pass
            time.sleep(0.001)

    def result(self, t0=None, t1=None):
        """median SM clock over the timed window [t0, t1]; if the window is too short
        for three samples the warm-up samples (same load) are included"""
        inside = [c for (t, c) in self.samples if t0 is None or t0 <= t <= t1]
        window = "timed"
        if len(inside) < 3:
            inside = [c for (_, c) in self.samples]
            window = "warmup+timed"
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": 0}
        s = sorted(inside)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(s), "window": window}


def cpu_reference_run(batch, steps, warmup, threads=0):
    """The CPU restatement of the reference's NTL path (oracle/cpu_ref.cpp) on the
    same workload; returns (shares/s, seconds per step, threads used)."""
    import numpy as np

    import __graft_entry__ as graft
    from oracle import cpu_ref
    from oracle import hbmpc_oracle as orc

    graft.build_oracle()
    ref = cpu_ref.CpuRef()
    pt = orc.EvalPoint(P, N_PARTIES, True)
    c = synth(batch, K, 0xB202)
    enc = ref.fft_batch_evaluate_limbs(c, pt.omega, P, pt.order, N_PARTIES, threads=threads)
    y = np.ascontiguousarray(enc[:, ZS, :])
    for _ in range(warmup):
        ref.fft_batch_evaluate_limbs(c, pt.omega, P, pt.order, N_PARTIES, threads=threads)
        ref.fft_batch_interpolate_limbs(ZS, y, pt.omega, P, pt.order, threads=threads)
    t0 = time.perf_counter()
    for _ in range(steps):
        ref.fft_batch_evaluate_limbs(c, pt.omega, P, pt.order, N_PARTIES, threads=threads)
        r = ref.fft_batch_interpolate_limbs(ZS, y, pt.omega, P, pt.order, threads=threads)
    dt = (time.perf_counter() - t0) / steps
    assert np.array_equal(r, c), "CPU reference round trip failed"
    used = threads if threads > 0 else ref.max_threads()
    return batch * K / dt, dt, used


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sample = 8192
    steps = max(1, min(args.steps, 40))
    warm = max(1, min(args.warmup, 3))
    value, dt, cores = cpu_reference_run(sample, steps, warm)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "warmup": warm, "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 mod p",
        "data": "synthetic",
        "config": {"workload": "n=16 t=5 NTT encode+interpolate (BASELINE configs[1])",
                   "batch_polys_per_step": sample, "shares_per_poly": K, "field": "BLS12-381 r"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} polynomials x {steps} steps; C++ restatement of "
                                   "rsdecode_impl.h (NTL itself is not installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from honeybadgermpc_b200 import _native
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.ntl import pack_vec
    from honeybadgermpc_b200.polynomial import EvalPoint
    from honeybadgermpc_b200.sharding import all_gather_rows

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    dev = torch.device("cuda", local)

    batch, sets = args.batch, args.sets
    pt = EvalPoint(GF(P), N_PARTIES, True)
    omega = pack_vec([pt.omega.value], P)[0]
    ctx = _native.Context(P, device=local)
    stream = torch.cuda.Stream(device=dev)
    ctx.set_stream(stream.cuda_stream)

    def dev_u64(a):
        return torch.from_numpy(a.view(np.int64)).to(dev)

    # ---- synthetic inputs, resident in HBM; `sets` rotating buffer sets so every
    # step reads cold data (sets * 71 MB >> 126 MB of L2)
    c, e, y, r = [], [], [], []
    zs_t = torch.tensor(ZS, device=dev)
    with torch.cuda.stream(stream):
        for s in range(sets):
            cs = dev_u64(synth(batch, K, 0xB202 + 977 * s + 131 * rank))
            es = torch.empty((batch, N_PARTIES, 4), dtype=torch.int64, device=dev)
            ctx.fft_batch_evaluate(omega, pt.order, cs.data_ptr(), batch, K, N_PARTIES,
                                   es.data_ptr(), _native.MEM_DEVICE)
            ys = es.index_select(1, zs_t).contiguous()
            es.zero_()
            c.append(cs)
            e.append(es)
            y.append(ys)
            r.append(torch.zeros((batch, K, 4), dtype=torch.int64, device=dev))
        depth = min(3, sets)
        gathered, handles, gather_mode = [], [], "none"
        if world > 1 and args.gather != "nccl":
            # symmetric memory: every rank's gather buffer mapped into every process, plus an
            # NVSwitch multicast address when the fabric offers one -> the interpolation kernel
            # stores its block straight into all ranks' buffers (fused compute + all-gather)
            try:
                import torch.distributed._symmetric_memory as symm_mem

                for _ in range(depth):
                    buf = symm_mem.empty((world * batch, K, 4), dtype=torch.int64, device=dev)
                    handles.append(symm_mem.rendezvous(buf, dist.group.WORLD))
                    gathered.append(buf)
                mc = int(getattr(handles[0], "multicast_ptr", 0) or 0)
                if args.gather == "copy":
                    gather_mode = "multimem-copy" if mc else "p2p-copy"
                elif args.gather == "auto":
                    gather_mode = "fused-multimem" if mc else "fused-p2p"
                else:
                    gather_mode = "fused-p2p"
            except Exception as exc:  # noqa: BLE001
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({exc!r}); using NCCL all-gather",
                          file=sys.stderr)
                gathered, handles = [], []
        if world > 1 and not handles:
            gather_mode = "nccl-overlapped"
            gathered = [torch.empty((world * batch, K, 4), dtype=torch.int64, device=dev)
                        for _ in range(depth)]
        pending = [None] * len(gathered)
        slot_done = [None] * len(gathered)
        side = torch.cuda.Stream(device=dev)
        # per-slot constants of the gather, marshalled once (the step loop is host-time critical)
        slot_peers = [_native.Context.peer_array(list(h.buffer_ptrs)) for h in handles]
        slot_mc = [int(getattr(h, "multicast_ptr", 0) or 0) for h in handles]
        written_ev = [torch.cuda.Event() for _ in handles]
        done_ev = [torch.cuda.Event() for _ in handles]
    zs32 = np.ascontiguousarray(ZS, dtype=np.int32)
    stream.synchronize()

    names = {"encode": ctx.last_kernel()}  # the set-up loop above ended with an encode
    c_ptr = [t.data_ptr() for t in c]
    e_ptr = [t.data_ptr() for t in e]
    y_ptr = [t.data_ptr() for t in y]
    r_ptr = [t.data_ptr() for t in r]

    # For N > 1 the all-gather of the decoded block is fused into the interpolation kernel (or runs
    # on a side stream, see below) and the independent encode of the step runs on a second stream:
    # the fused kernel leaves SM time free while it waits on NVLink.  Measured at 8 ranks with the
    # current kernels: 2.19e10 shares/s with the encode overlapped, 2.02e10 in order (with the
    # first kernels of this round it was the other way round: 1.91e10 against 2.18e10).
    # N = 1: the encode and the interpolation of a step are independent launches, so they go
    # to two streams: each kernel fills the GPU in a single wave (1024 CTAs for 1036 slots)
    # and spends ~8 us of its ~30 us ramping up and draining; on two streams the next
    # kernel's CTAs take over SM by SM as the previous kernel's CTAs retire.  The per-kernel
    # durations of the roofline come from a second, serial pass over the same steps.
    overlap_encode = (world > 1 and gather_mode.startswith("fused")) or (world == 1 and not args.serial)
    if args.overlap_encode != "auto":
        overlap_encode = args.overlap_encode == "on"
    enc_stream = torch.cuda.Stream(device=dev) if overlap_encode else stream

    # one context per stream (no hbg_ctx_set_stream inside the step loop: the loop is host-time
    # critical -- a step is ~40 us of GPU work)
    ctx_enc = _native.Context(P, device=local) if overlap_encode else ctx
    ctx_enc.set_stream(enc_stream.cuda_stream)

    def step(s, evs=None):
        if evs is not None:
            evs[0].record(enc_stream)
        ctx_enc.fft_batch_evaluate(omega, pt.order, c_ptr[s], batch, K, N_PARTIES, e_ptr[s], _native.MEM_DEVICE)
        if evs is not None:
            evs[1].record(enc_stream)
        if evs is not None and overlap_encode:
            evs[3].record(stream)
        fused = handles and gather_mode.startswith("fused")
        if handles:
            slot = s % len(gathered)
            h = handles[slot]
            if slot_done[slot] is not None:
                stream.wait_event(slot_done[slot])  # the previous gather into this slot is complete everywhere
        if fused:
            use_mc = slot_mc[slot] if gather_mode == "fused-multimem" else 0
            ctx.fft_batch_interpolate_allgather(omega, pt.order, zs32, y_ptr[s], batch,
                                                slot_peers[slot], use_mc, rank)
        else:
            ctx.fft_batch_interpolate(omega, pt.order, zs32, y_ptr[s], batch, r_ptr[s],
                                      _native.MEM_DEVICE)
        if evs is not None:
            evs[2].record(stream)
        if handles:
            # completion of the gather (= every rank's block has landed in every buffer) is a
            # device-side barrier over the symmetric-memory signal pads; it runs on a side
            # stream so this rank's next encode does not wait for the slowest rank
            slot = s % len(gathered)
            written = written_ev[slot]
            written.record(stream)
            side.wait_event(written)
            if not fused:
                # copy kernel: a few CTAs push this rank's block into every rank's buffer
                # (multimem.st -> replicated by the NVSwitch) while `stream` moves on
                ctx.set_stream(side.cuda_stream)
                ctx.allgather_block(r_ptr[s], batch * K * E, slot_peers[slot],
                                    slot_mc[slot] if gather_mode == "multimem-copy" else 0,
                                    rank * batch * K * E, args.gather_ctas)
                ctx.set_stream(stream.cuda_stream)
            handles[slot].barrier()  # on torch's current stream = `side` (set around the step loops)
            slot_done[slot] = done_ev[slot]
            slot_done[slot].record(side)
        elif world > 1:
            # the one collective of the path: reassemble the decoded blocks on every rank.
            # It runs on NCCL's stream and overlaps the next step's kernels; the buffer
            # pair (r[s], gathered[slot]) is only reused after its gather has completed.
            slot = s % len(gathered)
            if pending[slot] is not None:
                pending[slot].wait()
            _, pending[slot] = all_gather_rows(r[s], world * batch, out=gathered[slot], async_op=True)

    def barrier():
        for i, w in enumerate(pending):
            if w is not None:
                w.wait()
                pending[i] = None
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # torch's *current* stream only matters for the symmetric-memory barrier of the fused
    # gather (our kernels and events name their streams explicitly): make it `side` once
    # instead of entering a stream context on every step
    with torch.cuda.stream(side if handles else stream):
        sampler = ClockSampler(local)
        sampler.start()
        for i in range(args.warmup):
            step(i % sets)
            if i == 0:
                names["interpolate"] = ctx.last_kernel()
        barrier()
        # per-kernel events: every step when the kernels run in order on one stream; with
        # overlapped streams only every `--event-every`-th step (they cost host time, and for
        # N = 1 the roofline durations come from the serial pass anyway)
        ev_every = 1 if not overlap_encode else max(1, args.event_every)
        evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] if i % ev_every == 0 else None
               for i in range(args.steps)]
        t_start = torch.cuda.Event(enable_timing=True)
        t_end = torch.cuda.Event(enable_timing=True)
        launches0 = ctx.launch_count() + (ctx_enc.launch_count() if ctx_enc is not ctx else 0)
        barrier()
        host_t0 = time.perf_counter()
        t_start.record(stream)
        for i in range(args.steps):
            step((args.warmup + i) % sets, evs[i])
        stream.wait_stream(enc_stream)
        t_end.record(stream)
        host_enqueued = time.perf_counter()
        barrier()
        host_t1 = time.perf_counter()
        launches = ctx.launch_count() + (ctx_enc.launch_count() if ctx_enc is not ctx else 0) - launches0
        serial_evs = None
        if world == 1 and overlap_encode:
            # serial pass (not part of `value`): the same steps with both kernels on one stream,
            # so each kernel's CUDA-event duration is its own
            enc_stream_saved, enc_stream = enc_stream, stream
            ctx_enc.set_stream(stream.cuda_stream)
            serial_evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)]
                          for _ in range(min(args.steps, 200))]
            for i, ev in enumerate(serial_evs):
                step((args.warmup + i) % sets, ev)
            barrier()
            enc_stream = enc_stream_saved
            ctx_enc.set_stream(enc_stream.cuda_stream)
        sampler.stop_flag.set()
        sampler.join()
    total_ms = t_start.elapsed_time(t_end)
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    timed = [ev for ev in evs if ev is not None]
    enc_ms = sum(ev[0].elapsed_time(ev[1]) for ev in timed) / len(timed)
    dec_ms = sum(ev[3 if overlap_encode else 1].elapsed_time(ev[2]) for ev in timed) / len(timed)
    overlapped_ms = None
    if serial_evs is not None:
        overlapped_ms = {"encode": enc_ms, "interpolate": dec_ms}
        enc_ms = sum(ev[0].elapsed_time(ev[1]) for ev in serial_evs) / len(serial_evs)
        dec_ms = sum(ev[1].elapsed_time(ev[2]) for ev in serial_evs) / len(serial_evs)

    # parity inside the bench: every decoded block equals its coefficients
    used = min(sets, args.warmup + args.steps)
    if handles:
        last = (args.warmup + args.steps - 1) % sets
        g = gathered[last % len(gathered)]
        assert torch.equal(g[rank * batch:(rank + 1) * batch], c[last]), "gather: own block mismatch"
        sums = g.view(world, -1).sum(dim=1)
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "fused gather: ranks disagree on the gathered blocks"
    else:
        for s in range(used):
            assert torch.equal(r[s], c[s]), f"round trip mismatch in buffer set {s}"
        if world > 1:
            last = (args.warmup + args.steps - 1) % sets
            assert torch.equal(gathered[last % len(gathered)][rank * batch:(rank + 1) * batch], r[last])

    ms_per_step = total_ms / args.steps
    value = world * batch * K / (ms_per_step * 1e-3)

    # ---- roofline of the dominant kernel (algorithmic bytes)
    enc_bytes = batch * (K + N_PARTIES) * E
    dec_bytes = batch * (K + K) * E
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as fh:
            peaks = json.load(fh)
    except OSError:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if peaks else "fallback 6650 GB/s"
    if enc_ms >= dec_ms:
        dom, dom_ms, dom_bytes = "encode", enc_ms, enc_bytes
    else:
        dom, dom_ms, dom_bytes = "interpolate", dec_ms, dec_bytes
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            traffic = json.load(fh).get(names[dom])
    except (OSError, ValueError):
        pass
    # the binding resource is the 64-bit integer multiply-add pipe: algorithmic
    # IMAD.WIDE per polynomial against the measured pipe rate
    # (tools/microbench3.cu on this pool: carry chains of IMAD.WIDE at 31 per clock per SM =
    # 9.0e12 /s at 1965 MHz; one warp-wide IMAD.WIDE holds the fmaheavy pipe for 4 cycles)
    imad = {"encode": 15 * 103, "interpolate": K * K * 64 + K * 48}
    imad_peak = 9.0e12
    imad_rate = imad[dom] * batch / (dom_ms * 1e-3)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "kernel": f"{dom}: {names[dom]}",
                "peak_source": peak_src,
                "kernel_ms": {"encode": enc_ms, "interpolate": dec_ms}, "kernels": names,
                "kernel_ms_source": ("serial pass after the timed region (the timed steps overlap the two "
                                     "kernels on two streams)" if serial_evs is not None else "timed region"),
                "kernel_ms_overlapped": overlapped_ms,
                "step_GBps": (enc_bytes + dec_bytes) / (ms_per_step * 1e-3) / 1e9,
                "int_pipe": {"achieved_imad_wide_per_s": imad_rate, "peak_imad_wide_per_s": imad_peak,
                             "frac": imad_rate / imad_peak,
                             "imad_wide_per_polynomial": imad[dom]},
                "note": "256-bit modular arithmetic: the kernel is bound by the IMAD.WIDE pipe, not by "
                        "HBM (DESIGN.md section 4); both fractions are reported"}

    # ---- end to end: the C-ABI call a reference-side binding makes, HOST buffers
    # (pinned), H2D + kernel + D2H inside the timed region
    e2e_steps = max(3, min(args.steps, 60))
    hc = torch.from_numpy(synth(batch, K, 0xE2E + rank).view(np.int64)).pin_memory()
    he = torch.empty((batch, N_PARTIES, 4), dtype=torch.int64).pin_memory()
    hy = torch.empty((batch, K, 4), dtype=torch.int64).pin_memory()
    hr = torch.empty((batch, K, 4), dtype=torch.int64).pin_memory()

    def e2e_step():
        ctx.fft_batch_evaluate(omega, pt.order, hc.data_ptr(), batch, K, N_PARTIES, he.data_ptr(),
                               _native.MEM_HOST)
        ctx.fft_batch_interpolate(omega, pt.order, ZS, hy.data_ptr(), batch, hr.data_ptr(),
                                  _native.MEM_HOST)

    e2e_step()
    hy.copy_(he[:, ZS, :])
    # asynchronous host mode: the two calls of a step overlap on the PCIe link; every step
    # ends with hbg_ctx_synchronize, i.e. with its results readable in host memory
    ctx.set_host_async(True)
    for _ in range(2):
        e2e_step()
        ctx.synchronize()
    barrier()
    t0 = time.perf_counter()
    marks = [t0]
    for _ in range(e2e_steps):
        e2e_step()
        ctx.synchronize()
        marks.append(time.perf_counter())
    torch.cuda.synchronize()
    e2e_dt = (time.perf_counter() - t0) / e2e_steps
    per_step = sorted(b - a for a, b in zip(marks, marks[1:]))
    ctx.set_host_async(False)
    assert torch.equal(hr, hc), "end-to-end round trip mismatch"
    if world > 1:
        tt = torch.tensor([e2e_dt], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e_dt = float(tt.item())
    e2e = {"value": world * batch * K / e2e_dt, "unit": UNIT,
           "h2d_bytes_per_step": 2 * batch * K * E,
           "d2h_bytes_per_step": batch * (N_PARTIES + K) * E,
           "ms_per_step": e2e_dt * 1e3,
           "ms_per_step_min_median_max": [per_step[0] * 1e3, per_step[len(per_step) // 2] * 1e3,
                                          per_step[-1] * 1e3],
           "boundary": "hbg_fft_batch_evaluate + hbg_fft_batch_interpolate, HBG_MEM_HOST, pinned buffers, "
                       "host_async on, hbg_ctx_synchronize at the end of every step"}

    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            sample = 16384
            v, dt, cores = cpu_reference_run(sample, 40, 1)
            cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                   "sample": f"{sample} polynomials x 40 steps of the same encode+interpolate; "
                             "oracle/cpu_ref.cpp (C++ restatement of rsdecode_impl.h, OpenMP over the batch)",
                   "ms_per_step": dt * 1e3}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u32x8 Montgomery (256-bit integer mod p)",
            "data": "synthetic",
            "config": {"workload": "n=16 t=5 NTT encode+interpolate (BASELINE configs[1])",
                       "batch_polys_per_gpu": batch, "shares_per_poly": K,
                       "field": "BLS12-381 r", "z": ZS,
                       "l2": f"{sets} rotating buffer sets of {(enc_bytes + dec_bytes) / 1e6:.0f} MB "
                             "(inputs+outputs larger than the 126 MB L2)",
                       "streams": ("encode and interpolate of a step on two streams" if overlap_encode
                                   else "one stream"),
                       "parallelism": f"batch shard x{world}" + (
                           f" + all-gather ({gather_mode})"
                           + (", encode overlapped on a second stream" if overlap_encode else "")
                           if world > 1 else ""),
                       "polys_per_s": value / K},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches,
            "host_enqueue_ms_per_step": (host_enqueued - host_t0) * 1e3 / args.steps,
            "clocks": sampler.result(host_t0, host_t1),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return line


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=400)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=65536)
    ap.add_argument("--sets", type=int, default=6)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--overlap-encode", default="auto", choices=["auto", "on", "off"],
                    help="run the encode of a step on its own stream (auto: N = 1, and N > 1 with the fused gather)")
    ap.add_argument("--event-every", type=int, default=8,
                    help="with overlapped streams: record per-kernel events on every n-th step only")
    ap.add_argument("--serial", action="store_true",
                    help="N=1: run the two kernels of a step back to back on one stream")
    ap.add_argument("--gather", default="auto", choices=["auto", "p2p", "copy", "nccl"],
                    help="N>1: auto = the interpolation kernel stores its block into every rank's "
                         "symmetric-memory buffer itself (multimem.st when a multicast address exists); "
                         "p2p = the same with peer stores only; copy = separate side-stream copy kernel; "
                         "nccl = overlapped NCCL all-gather")
    ap.add_argument("--gather-ctas", type=int, default=16)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
