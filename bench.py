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


def choose_sm_split(tiles, n_sm, ratio=2.4):
    """SMs for the encode launches when the two kernels of a step run side by side (the interpolation
    gets the rest): the split whose longer chain is shortest, in tile rounds weighted by the per-tile
    cost ratio encode : interpolate (16 vs 6 outputs per row; 2.4 from the sweep in profiles/), and
    among equals the one closest to the work-proportional split.  65 536 rows on 148 SMs: 104 + 44
    (5 and 12 rounds)."""
    ideal = n_sm * ratio / (1.0 + ratio)
    best = None
    for e_sm in range(n_sm // 2, n_sm - 8):
        enc_t = -(-tiles // e_sm) * ratio
        dec_t = -(-tiles // (n_sm - e_sm)) * 1.0
        key = (max(enc_t, dec_t), abs(e_sm - ideal))
        if best is None or key < best[0]:
            best = (key, e_sm)
    return best[1]


def synth(batch, width, seed):
    """uniform in [0, 2^254) subset of [0, p): canonical residues as uint64[batch,width,4]"""
    import numpy as np

    rng = np.random.default_rng(seed)
    a = rng.integers(0, 2 ** 63, size=(batch, width, 4), dtype=np.uint64) * np.uint64(2) + \
        rng.integers(0, 2, size=(batch, width, 4), dtype=np.uint64)
    a[:, :, 3] >>= np.uint64(2)
    return a


def gpu_numa_cpus(props):
    """(NUMA node, its CPUs) of the GPU with these device properties, from sysfs; (None, None)
    when the platform does not say (single node, container without /sys, node -1)."""
    try:
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{bdf}/numa_node") as fh:
            node = int(fh.read().strip())
        if node < 0:
            return None, None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as fh:
            cpus = parse_cpulist(fh.read())
        return node, (cpus or None)
    except (OSError, ValueError, AttributeError):
        return None, None


def parse_cpulist(text):
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11}"""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


class numa_local:
    """Run a block with the calling thread confined to the CPUs of the GPU's NUMA node: pinned host
    buffers allocated inside are first touched there, and the thread that enqueues the copies stays
    next to them.  No-op where the node is unknown.  The previous affinity is restored on exit (the
    CPU baseline counts its threads from the affinity mask)."""

    def __init__(self, props):
        self.node, cpus = gpu_numa_cpus(props)
        self.saved = None
        try:
            allowed = os.sched_getaffinity(0)
        except (AttributeError, OSError):
            allowed = None
        self.cpus = (cpus & allowed) if (cpus and allowed) else None
        self.applied = False

    def __enter__(self):
        if self.cpus:
            try:
                self.saved = os.sched_getaffinity(0)
                os.sched_setaffinity(0, self.cpus)
                self.applied = True
            except OSError:
                self.applied = False
        return self

    def __exit__(self, *exc):
        if self.applied and self.saved:
            try:
                os.sched_setaffinity(0, self.saved)
            except OSError:
                pass
        return False

    def describe(self):
        return {"gpu_numa_node": self.node, "cpus_used": len(self.cpus) if self.applied else None,
                "applied": self.applied}


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


def host_threads():
    """every core this process may run on -- torchrun exports OMP_NUM_THREADS=1, which must
    not shrink the CPU baseline"""
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_reference_run(batch, steps, warmup, threads=0):
    """The CPU restatement of the reference's NTL path (oracle/cpu_ref.cpp) on the
    same workload; returns (shares/s, seconds per step, threads used)."""
    import numpy as np

    import __graft_entry__ as graft
    from oracle import cpu_ref
    from oracle import hbmpc_oracle as orc

    graft.build_oracle()
    ref = cpu_ref.CpuRef()
    threads = threads if threads > 0 else host_threads()
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
    return batch * K / dt, dt, threads


def run_reference(args):
    """The reference arm: the CPU implementation of the path (the C++ restatement of
    rsdecode_impl.h -- NTL itself cannot be installed here) on this box's host cores, same
    workload, batch and warm-up as the B200 arm; the step count is bounded so the run
    stays within about a minute."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    batch = args.batch
    probe_v, probe_dt, cores = cpu_reference_run(batch, 1, 1)
    steps = max(3, min(args.steps, int(45.0 / probe_dt)))
    value, dt, cores = cpu_reference_run(batch, steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
        "n_gpus": args.gpus, "steps": steps, "steps_requested": args.steps, "warmup": args.warmup,
        "ms_per_step": dt * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64x4 mod p",
        "data": "synthetic",
        "config": {"workload": "n=16 t=5 NTT encode+interpolate (BASELINE configs[1])",
                   "batch_polys_per_step": batch, "shares_per_poly": K, "field": "BLS12-381 r", "z": ZS},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{batch} polynomials x {steps} steps on {cores} threads (explicit, "
                                   "OMP_NUM_THREADS ignored); C++ restatement of rsdecode_impl.h "
                                   "(NTL itself is not installable here)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def run_cfg5(args, world, rank, local, dev):
    """BASELINE configs[4] as a STRONG-scaling record: n = 128, t = 42; 1 048 576 polynomials in
    total, the batch axis sharded over the ranks; one pass = NTT-128 encode of this rank's shard
    + interpolation from 43 scattered shares + all-gather, so that every rank ends up with all
    2^20 x 43 opened values.  The shard is processed in pieces so the gather of one piece
    overlaps the kernels of the next.  Returns the dict stored under "cfg5_strong"."""
    import random

    import numpy as np
    import torch
    import torch.distributed as dist

    from honeybadgermpc_b200 import _native
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.ntl import pack_vec
    from honeybadgermpc_b200.polynomial import EvalPoint
    from honeybadgermpc_b200.sharding import ShardedReconstructor

    n, k = 128, 43
    total = args.cfg5_polys
    shard = total // world
    piece = min(shard, args.cfg5_piece)
    parts = shard // piece
    pt = EvalPoint(GF(P), n, True)
    omega = pack_vec([pt.omega.value], P)[0]
    zs = sorted(random.Random(5).sample(range(n), k))
    gather = "auto" if args.gather.startswith("fused") else args.gather  # k = 43: no fused epilogue
    if gather == "auto" and world >= 3:
        # cfg5 is bound by its kernels, not by the gather (7x more arithmetic per gathered byte): the
        # multicast copy kernel, which needs no SMs set aside, beats the bulk-copy default of the
        # reconstructor at 4 ranks (2.47 against 3.17 ms per pass)
        gather = "mc"
    rec = ShardedReconstructor(P, omega, pt.order, zs, piece, device=local, depth=2, gather=gather,
                               copy_ctas=args.gather_ctas, parts=parts)
    ctx, stream = rec.ctx, rec.stream
    enc_stream = torch.cuda.Stream(device=dev)
    ctx_enc = _native.Context(P, device=local)
    ctx_enc.set_stream(enc_stream.cuda_stream)
    if rec.compute_sms > 0:
        ctx_enc.set_sm_limit(rec.compute_sms)
    rng = np.random.default_rng(0xB205 + rank)
    zs_t = torch.tensor(zs, device=dev)
    with torch.cuda.stream(stream):
        c = torch.from_numpy(rng.integers(0, 2 ** 62, size=(shard, k, 4), dtype=np.uint64).view(np.int64)).to(dev)
        e = torch.empty((shard, n, 4), dtype=torch.int64, device=dev)
        ctx.fft_batch_evaluate(omega, pt.order, c.data_ptr(), shard, k, n, e.data_ptr(), _native.MEM_DEVICE)
        y = e.index_select(1, zs_t).contiguous()
        e.zero_()
    stream.synchronize()
    names = {"encode": ctx.last_kernel()}
    pb_c, pb_e, pb_y = piece * k * E, piece * n * E, piece * k * E

    def one_pass(slot):
        enc_stream.wait_stream(stream)
        for p in range(parts):
            ctx_enc.fft_batch_evaluate(omega, pt.order, c.data_ptr() + p * pb_c, piece, k, n,
                                       e.data_ptr() + p * pb_e, _native.MEM_DEVICE)
            rec.open(y.data_ptr() + p * pb_y, slot=slot, part=p)
        stream.wait_stream(enc_stream)
        rec.finish(slot)
        if rec.handles or rec.world == 1:
            stream.wait_event(rec.done_ev[slot])  # the pass ends when every rank's pieces have landed
        else:
            with torch.cuda.stream(stream):
                rec.wait(slot)

    def barrier():
        rec.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for w in range(2):
        one_pass(w % 2)
    names["interpolate"] = ctx.last_kernel()
    barrier()
    passes = args.cfg5_passes
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record(stream)
    for i in range(passes):
        one_pass(i % 2)
    t1.record(stream)
    barrier()
    ms = t0.elapsed_time(t1) / passes
    if world > 1:
        tt = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    g = rec.gathered[(passes - 1) % 2]
    assert torch.equal(g[rank * shard:(rank + 1) * shard], c), "cfg5: decoded shard != coefficients"
    assert torch.equal(e.index_select(1, zs_t), y), "cfg5: encode output is wrong"
    if world > 1:
        sums = g.view(world, -1).sum(dim=1)
        lo, hi = sums.clone(), sums.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        assert torch.equal(lo, hi), "cfg5: ranks disagree on the gathered result"
    # per-kernel time of one piece (serial, events)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
    ctx_enc.set_stream(stream.cuda_stream)
    ev[0].record(stream)
    ctx_enc.fft_batch_evaluate(omega, pt.order, c.data_ptr(), piece, k, n, e.data_ptr(), _native.MEM_DEVICE)
    ev[1].record(stream)
    ctx.fft_batch_interpolate(omega, pt.order, np.ascontiguousarray(zs, dtype=np.int32), y.data_ptr(), piece,
                              rec.own_block_ptr(0, 0), _native.MEM_DEVICE)
    ev[2].record(stream)
    barrier()
    ingress = (world - 1) * shard * k * E
    out = {"workload": "n=128 t=42 NTT-128 encode + interpolate from 43 scattered shares + all-gather "
                       "(BASELINE configs[4]), STRONG scaling: the total is fixed",
           "total_polys": total, "polys_per_gpu": shard, "piece": piece, "scaling": "strong",
           "ms_per_pass": ms, "value": total * k / (ms * 1e-3), "unit": UNIT, "passes": passes,
           "gather": rec.mode, "kernels": names,
           "kernel_ms_per_piece": {"encode": ev[0].elapsed_time(ev[1]), "interpolate": ev[1].elapsed_time(ev[2])},
           "nvlink_ingress_bytes_per_pass": ingress,
           "nvlink_floor_ms": ingress / 770e9 * 1e3,
           "algorithmic_GBps": total / world * (3 * k + n) * E / (ms * 1e-3) / 1e9}
    del rec, c, e, y
    torch.cuda.empty_cache()
    return out


def run_b200(args):
    import numpy as np
    import torch
    import torch.distributed as dist

    from honeybadgermpc_b200 import _native
    from honeybadgermpc_b200.field import GF
    from honeybadgermpc_b200.ntl import pack_vec
    from honeybadgermpc_b200.polynomial import EvalPoint
    from honeybadgermpc_b200.sharding import ShardedReconstructor

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
    # the sharded reconstructor owns the interpolation context / stream and the gather slots
    depth = sets if world == 1 else min(3, sets)
    rec = ShardedReconstructor(P, omega, pt.order, ZS, batch, device=local, depth=depth,
                               gather=args.gather, copy_ctas=args.gather_ctas, compute_sms=args.sm_limit)
    ctx, stream = rec.ctx, rec.stream
    if args.matvec_path != "auto":
        ctx.set_matvec_path(args.matvec_path)

    def dev_u64(a):
        return torch.from_numpy(a.view(np.int64)).to(dev)

    # ---- synthetic inputs, resident in HBM; `sets` rotating buffer sets so every
    # step reads cold data (sets * 71 MB >> 126 MB of L2)
    c, e, y = [], [], []
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
    stream.synchronize()
    names = {"encode": ctx.last_kernel()}  # the set-up loop above ended with an encode
    c_ptr = [t.data_ptr() for t in c]
    e_ptr = [t.data_ptr() for t in e]
    y_ptr = [t.data_ptr() for t in y]

    # The encode and the interpolation of a step are independent launches: they go to two
    # streams, so one kernel's CTAs take over SM by SM as the other's retire and neither's
    # fill / drain phase is exposed.  The per-kernel durations of the roofline come from a
    # second, serial pass over the same steps.
    overlap_encode = not args.serial
    if args.overlap_encode != "auto":
        overlap_encode = args.overlap_encode == "on"
    enc_stream = torch.cuda.Stream(device=dev) if overlap_encode else stream
    ctx_enc = _native.Context(P, device=local) if overlap_encode else ctx
    ctx_enc.set_stream(enc_stream.cuda_stream)
    if args.matvec_path != "auto":
        ctx_enc.set_matvec_path(args.matvec_path)
    if rec.compute_sms > 0:  # --sm-limit, or the reconstructor's own choice (copy kernel on its own SMs)
        ctx_enc.set_sm_limit(rec.compute_sms)
    if args.tc_store != "direct":
        ctx.set_tc_store(args.tc_store)
        ctx_enc.set_tc_store(args.tc_store)
    sm_split = None
    want_split = args.sm_split
    if want_split < 0:  # auto: N = 1, and the copy-engine gather (a copy KERNEL's SMs would come first)
        want_split = 0
        if ((world == 1 or rec.mode == "ce-copy-signal") and overlap_encode
                and args.matvec_path in ("auto", "tc") and rec.compute_sms == 0):
            # Persistent tensor-core CTAs own a whole SM, so two full-GPU launches run one after the
            # other and every launch pays its own pipeline fill on every SM.  Side by side on
            # disjoint SMs the fills overlap.  Split = argmin over e of the longer chain, in tile
            # rounds weighted by the per-tile cost ratio (measured); among equals the one closest to
            # the work-proportional split.  65 536 rows: 104 + 44 SMs (5 and 12 rounds).
            n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
            want_split = choose_sm_split((batch + 127) // 128, n_sm)
    if overlap_encode and want_split > 0:
        # the two kernels side by side on disjoint SMs instead of one after the other
        n_sm = torch.cuda.get_device_properties(dev).multi_processor_count
        sm_split = [min(want_split, n_sm - 1), n_sm - min(want_split, n_sm - 1)]
        ctx_enc.set_sm_limit(sm_split[0])
        ctx.set_sm_limit(sm_split[1])

    def encode(s):
        ctx_enc.fft_batch_evaluate(omega, pt.order, c_ptr[s], batch, K, N_PARTIES, e_ptr[s], _native.MEM_DEVICE)

    def step(i, evs=None):
        s = i % sets
        if evs is not None:
            evs[0].record(enc_stream)
        encode(s)
        if evs is not None:
            evs[1].record(enc_stream)
            if overlap_encode:
                evs[3].record(stream)
        rec.open(y_ptr[s], slot=i % depth)
        if evs is not None:
            evs[2].record(stream)
        rec.finish(i % depth)  # no reader in the benchmark: the slot is handed back at once

    def barrier():
        rec.drain()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local)
    sampler.start()
    for i in range(max(args.warmup, sets)):
        step(i)
        if i == 0:
            names["interpolate"] = ctx.last_kernel()
    barrier()

    # ---- the timed region: `unit` = lcm(sets, depth) consecutive steps captured into ONE CUDA
    # graph (the Python step loop costs ~20 us per step, more than the GPU work of a step),
    # replayed until at least --steps steps AND --min-ms of device time have run
    unit = int(sets * depth // np.gcd(sets, depth)) * args.graph_units
    graph = None
    if not args.no_graph:
        try:
            graph = rec.capture(
                [y_ptr[i % sets] for i in range(unit)],
                begin=(lambda: enc_stream.wait_stream(stream)) if overlap_encode else None,
                extra=lambda i: encode(i % sets),
                finish=(lambda: stream.wait_stream(enc_stream)) if overlap_encode else None)
        except Exception as exc:  # noqa: BLE001 - e.g. a collective that cannot be captured
            if rank == 0:
                print(f"[bench] CUDA graph capture failed ({exc!r}); eager step loop", file=sys.stderr)
            graph = None
            barrier()
    launches_per_step = 2 + ((4 if rec.mode in ("ce-copy-signal",) or rec.mode.startswith("fused") else 3)
                             if rec.signal else 1 if rec.mode.endswith("copy") else 0)

    def run_steps(n_steps):
        """enqueue n_steps (a multiple of `unit` when the graph is used)"""
        if graph is not None:
            with torch.cuda.stream(stream):
                for _ in range(n_steps // unit):
                    graph.replay()
        else:
            for i in range(n_steps):
                step(i)
            stream.wait_stream(enc_stream)

    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    # calibration: how many steps make --min-ms
    barrier()
    t_start.record(stream)
    run_steps(unit * 4)
    t_end.record(stream)
    barrier()
    est_ms = t_start.elapsed_time(t_end) / (unit * 4)
    if world > 1:
        tt = torch.tensor([est_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        est_ms = float(tt.item())
    steps = max(args.steps, int(np.ceil(args.min_ms / est_ms)))
    steps = int(np.ceil(steps / unit)) * unit
    barrier()
    host_t0 = time.perf_counter()
    t_start.record(stream)
    run_steps(steps)
    t_end.record(stream)
    host_enqueued = time.perf_counter()
    barrier()
    host_t1 = time.perf_counter()
    total_ms = t_start.elapsed_time(t_end)
    if world > 1:
        tt = torch.tensor([total_ms], device=dev, dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        total_ms = float(tt.item())
    sampler.stop_flag.set()
    sampler.join()

    # ---- parity inside the bench, on what the timed steps left behind: every buffer set's
    # encode output restricted to z equals the interpolation input, the gathered blocks equal
    # the coefficients on every rank
    for s in range(sets):
        assert torch.equal(e[s].index_select(1, zs_t), y[s]), f"encode output of buffer set {s} is wrong"
    last = steps - 1
    for back in range(min(depth, steps)):
        i = last - back
        g = rec.gathered[i % depth]
        assert torch.equal(g[rank * batch:(rank + 1) * batch], c[i % sets]), \
            f"step {i}: decoded block != coefficients"
        if world > 1:
            sums = g.view(world, -1).sum(dim=1)
            lo, hi = sums.clone(), sums.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX)
            assert torch.equal(lo, hi), "all-gather: ranks disagree on the gathered blocks"

    # ---- serial pass (not part of `value`): the same steps, eager, both kernels on one stream,
    # so each kernel's CUDA-event duration is its own
    enc_saved = enc_stream
    enc_stream = stream
    ctx_enc.set_stream(stream.cuda_stream)
    n_serial = min(max(args.steps, 20), 200)
    serial_evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(n_serial)]
    overlap_saved, overlap_encode = overlap_encode, False
    for i, ev in enumerate(serial_evs):
        step(i, ev)
    barrier()
    overlap_encode = overlap_saved
    enc_stream = enc_saved
    ctx_enc.set_stream(enc_stream.cuda_stream)
    eager_ms = {"encode": sum(ev[0].elapsed_time(ev[1]) for ev in serial_evs) / n_serial,
                "interpolate": sum(ev[1].elapsed_time(ev[2]) for ev in serial_evs) / n_serial}
    enc_ms, dec_ms = eager_ms["encode"], eager_ms["interpolate"]
    kernel_src = ("serial eager pass after the timed region (the timed steps overlap the two kernels on two "
                  "streams inside a CUDA graph)")
    # The eager pass brackets every launch with events the host enqueues one by one: for kernels of
    # 10 us the gaps between a recorded event and the next launch reaching the GPU are inside the
    # interval.  The per-kernel durations the roofline uses therefore come from two more graphs, one
    # kernel each, `unit` launches back to back on the bench stream, timed as a whole.
    full_ms = {}
    if not args.no_graph:
        try:
            scratch = torch.empty_like(c[0])

            def chain_ms(fn):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=stream, capture_error_mode="thread_local"):
                    for i in range(unit):
                        fn(i % sets)
                a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(stream):
                    g.replay()
                    a0.record(stream)
                    for _ in range(4):
                        g.replay()
                    a1.record(stream)
                stream.synchronize()
                return a0.elapsed_time(a1) / (4 * unit)

            ctx_enc.set_stream(stream.cuda_stream)
            enc_ms = chain_ms(encode)
            ctx_enc.set_stream(enc_stream.cuda_stream)
            dec_ms = chain_ms(lambda s_: ctx.fft_batch_interpolate(omega, pt.order, ZS, y_ptr[s_], batch,
                                                                   scratch.data_ptr(), _native.MEM_DEVICE))
            assert torch.equal(scratch, c[(unit - 1) % sets]), "interpolation chain: wrong result"
            if sm_split:
                # the same two chains with every SM available: the kernels as such, next to the kernels
                # as the timed region runs them (each on its share of the SMs)
                ctx_enc.set_sm_limit(0)
                ctx.set_sm_limit(0)
                ctx_enc.set_stream(stream.cuda_stream)
                full_ms["encode"] = chain_ms(encode)
                ctx_enc.set_stream(enc_stream.cuda_stream)
                full_ms["interpolate"] = chain_ms(
                    lambda s_: ctx.fft_batch_interpolate(omega, pt.order, ZS, y_ptr[s_], batch, scratch.data_ptr(),
                                                         _native.MEM_DEVICE))
                ctx_enc.set_sm_limit(sm_split[0])
                ctx.set_sm_limit(sm_split[1])
            kernel_src = (f"two CUDA graphs of {unit} back-to-back launches of one kernel each on the bench stream, "
                          "rotating buffer sets, CUDA events around 4 replays (launch gaps inside a chain "
                          "included); `kernel_ms_eager` = the per-launch event brackets of an eager pass")
        except Exception as exc:  # noqa: BLE001 - keep the eager figures
            if rank == 0:
                print(f"[bench] per-kernel graphs failed ({exc!r}); eager per-launch events", file=sys.stderr)
            ctx_enc.set_stream(enc_stream.cuda_stream)
            barrier()

    ms_per_step = total_ms / steps
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
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as fh:
            tj = json.load(fh)
        traffic = tj.get(f"{dom}:{names[dom]}")
        traffic_src = tj.get("source")
    except (OSError, ValueError):
        pass
    # u8 multiply-accumulates of the tensor-core kernel (K = 32 d bytes per row, 32 columns per
    # output) against the nominal dense 8-bit rate
    macs = {"encode": N_PARTIES * 32 * K * 32, "interpolate": K * 32 * K * 32}
    nb_link = (world - 1) * batch * K * E  # bytes every rank must RECEIVE per step
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "traffic_source": traffic_src or "none committed for this kernel",
                "kernel": f"{dom}: {names[dom]}", "peak_source": peak_src,
                "kernel_ms": {"encode": enc_ms, "interpolate": dec_ms}, "kernels": names,
                "kernel_ms_source": kernel_src, "kernel_ms_eager": eager_ms,
                "kernel_ms_all_sms": full_ms or None,
                "frac_all_sms": (dom_bytes / (full_ms[dom] * 1e-3) / 1e9 / peak) if full_ms else None,
                "step_GBps": (enc_bytes + dec_bytes) / (ms_per_step * 1e-3) / 1e9,
                "step_frac": (enc_bytes + dec_bytes) / (ms_per_step * 1e-3) / 1e9 / peak,
                "tensor": {"u8_mac_per_s": macs[dom] * batch / (dom_ms * 1e-3),
                           "nominal_u8_mac_per_s": 2.25e15,
                           "note": "exact u8 x u8 -> s32 GEMM on tcgen05 (kind::i8); the kernel is bound "
                                   "by HBM / the TMEM-read epilogue, not by the MMA rate"},
                "note": "256-bit modular arithmetic as an integer GEMM whose constant operand absorbs "
                        "the reduction (DESIGN.md section 4)"}
    # ---- the write side of the roofline, measured live: a pure write stream (memset) reaches only
    # about half of the copy bandwidth on this part, and 65 % of a step's traffic is writes (the
    # encode: 73 %), so max(bytes / copy peak, written bytes / write peak) is the tighter floor
    wbuf = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
    w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        wbuf.zero_()
        w0.record(stream)
        for _ in range(5):
            wbuf.zero_()
        w1.record(stream)
    stream.synchronize()
    write_peak = 5 * wbuf.numel() / (w0.elapsed_time(w1) * 1e-3) / 1e9
    del wbuf
    written = {"encode": batch * N_PARTIES * E, "interpolate": batch * K * E}
    step_written = written["encode"] + written["interpolate"] + (nb_link if world > 1 else 0)
    floor_ms = {k: max(written[k] / write_peak, b / peak) / 1e6
                for k, b in (("encode", enc_bytes), ("interpolate", dec_bytes))}
    step_floor_ms = max(step_written / write_peak, (enc_bytes + dec_bytes + (2 * nb_link if world > 1 else 0)) / peak) / 1e6
    roofline["hbm_write"] = {
        "write_peak_GBps": write_peak, "how": "5 x memset of 1 GiB on the bench stream, CUDA events",
        "written_bytes_per_launch": written[dom], "kernel_floor_ms": floor_ms[dom],
        "kernel_frac_of_floor": floor_ms[dom] / dom_ms,
        "step_written_bytes": step_written, "step_floor_ms": step_floor_ms,
        "step_frac_of_floor": step_floor_ms / ms_per_step,
        "note": "floor = max(all bytes / copy peak, written bytes / write peak); `frac` above keeps the "
                "contract's definition (algorithmic bytes / copy peak)"}
    if world > 1:
        link = 770.0  # GB/s per direction per GPU, measured peer copy (B200_PROFILING.md)
        roofline["nvlink"] = {"ingress_bytes_per_step": nb_link, "link_GBps": link,
                              "floor_ms_per_step": nb_link / link / 1e6,
                              "achieved_GBps": nb_link / (ms_per_step * 1e-3) / 1e9,
                              "frac": nb_link / (ms_per_step * 1e-3) / 1e9 / link,
                              "note": "every rank receives the other ranks' decoded blocks: for N >= 4 the "
                                      "step is bound by NVLink ingress, not by the kernels"}

    # ---- end to end: the C-ABI call a reference-side binding makes, HOST buffers
    # (pinned), H2D + kernel + D2H inside the timed region
    # the pinned buffers and the enqueueing thread of the end-to-end leg live on the GPU's NUMA node
    # (8 ranks on one host: every rank's 71 MB per step crosses the socket interconnect otherwise)
    numa = numa_local(torch.cuda.get_device_properties(dev))
    with numa:
        e2e_steps = max(4, min(args.steps, 60))
        hc = torch.from_numpy(synth(batch, K, 0xE2E + rank).view(np.int64)).pin_memory()
        hy = torch.empty((batch, K, 4), dtype=torch.int64).pin_memory()
        # two sets of output buffers: the results of step i are waited for (hbg_ctx_wait_pending) while the
        # transfers of step i+1 are in flight, and a set is only written again after it has been waited for
        he = [torch.empty((batch, N_PARTIES, 4), dtype=torch.int64).pin_memory() for _ in range(2)]
        hr = [torch.empty((batch, K, 4), dtype=torch.int64).pin_memory() for _ in range(2)]

        def e2e_step(b):
            ctx.fft_batch_evaluate(omega, pt.order, hc.data_ptr(), batch, K, N_PARTIES, he[b].data_ptr(),
                                   _native.MEM_HOST)
            ctx.fft_batch_interpolate(omega, pt.order, ZS, hy.data_ptr(), batch, hr[b].data_ptr(),
                                      _native.MEM_HOST)

        e2e_limit = ctx_enc is not ctx and sm_split is not None
        if e2e_limit:
            ctx.set_sm_limit(0)  # one context, one stream: its launches take every SM
        e2e_step(0)
        hy.copy_(he[0][:, ZS, :])
        # asynchronous host mode: the two calls of a step overlap on the PCIe link, and so do consecutive
        # steps (H2D of the next under the D2H of the current: the D2H side, 46 of the 71 MB, is the floor)
        ctx.set_host_async(True)
        for i in range(2):
            e2e_step(i)
        ctx.synchronize()
        for b in range(2):
            he[b].zero_()
            hr[b].zero_()
        barrier()
        t0 = time.perf_counter()
        marks = [t0]
        for i in range(e2e_steps):
            e2e_step(i & 1)
            ctx.wait_pending(2)  # step i-1 (two calls) is complete in host memory
            marks.append(time.perf_counter())
        ctx.synchronize()
        torch.cuda.synchronize()
        e2e_dt = (time.perf_counter() - t0) / e2e_steps
        per_step = sorted(b - a for a, b in zip(marks, marks[1:]))
        ctx.set_host_async(False)
        if e2e_limit:
            ctx.set_sm_limit(sm_split[1])
        for b in range(2):
            assert torch.equal(hr[b], hc), "end-to-end round trip mismatch"
            assert torch.equal(he[b][:, ZS, :], hy), "end-to-end encode mismatch"
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
           "numa": numa.describe(),
           "boundary": "hbg_fft_batch_evaluate + hbg_fft_batch_interpolate, HBG_MEM_HOST, pinned buffers, "
                       "host_async on; after enqueueing step i the host waits for step i-1 "
                       "(hbg_ctx_wait_pending(ctx, 2)): every step's outputs are complete in host memory "
                       "before its buffers are reused, two output buffer sets"}

    cfg5 = None
    if args.cfg5 == "on" or (args.cfg5 == "auto" and args.batch == 65536):
        try:
            cfg5 = run_cfg5(args, world, rank, local, dev)
        except Exception as exc:  # noqa: BLE001 - the headline line must still be printed
            cfg5 = {"error": repr(exc)}
            if world > 1:
                raise

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
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
            "steps_requested": args.steps,
            "warmup": max(args.warmup, sets), "ms_per_step": ms_per_step, "timed_region_ms": total_ms,
            "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "u8 x u8 -> s32 tensor-core GEMM + 256-bit "
                                                             "integer reduction mod p (exact)",
            "data": "synthetic",
            "config": {"workload": "n=16 t=5 NTT encode+interpolate (BASELINE configs[1])",
                       "batch_polys_per_gpu": batch, "shares_per_poly": K,
                       "field": "BLS12-381 r", "z": ZS,
                       "l2": f"{sets} rotating buffer sets of {(enc_bytes + dec_bytes) / 1e6:.0f} MB "
                             "(inputs+outputs larger than the 126 MB L2)",
                       "streams": ("encode and interpolate of a step on two streams" if overlap_encode
                                   else "one stream"),
                       "sm_split": ({"encode": sm_split[0], "interpolate": sm_split[1]} if sm_split
                                    else f"every launch may use {rec.compute_sms} SMs" if rec.compute_sms
                                    else "every launch may use all SMs"),
                       "step_loop": (f"CUDA graph of {unit} steps, replayed" if graph is not None
                                     else "eager Python loop"),
                       "parallelism": f"batch shard x{world}" + (f" + all-gather ({rec.mode})" if world > 1 else ""),
                       "polys_per_s": value / K},
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "cfg5_strong": cfg5,
            "gpu_launches": launches_per_step * steps,
            "host_enqueue_ms_per_step": (host_enqueued - host_t0) * 1e3 / steps,
            "clocks": sampler.result(host_t0, host_t1),
        }
        print(json.dumps(line, default=lambda o: o.item() if hasattr(o, "item") else str(o)), flush=True)
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
    ap.add_argument("--gather", default="auto", choices=["auto", "ce", "mc", "p2p", "bulk", "fused", "fused-barrier", "copy", "nccl"],
                    help="N>1: auto = ce = local store + copy-engine peer copies on a side stream, slot hand-over "
                         "by device flags; mc = multimem.st copy kernel + flags; p2p = peer-store copy kernel "
                         "+ flags; "
                         "fused = the kernel epilogue stores into every rank's buffer, two symmetric-memory "
                         "barriers per step; copy = copy kernel + barriers; nccl = overlapped NCCL all-gather")
    ap.add_argument("--gather-ctas", type=int, default=64)
    ap.add_argument("--cfg5", default="auto", choices=["auto", "on", "off"],
                    help="also run BASELINE configs[4] (n=128, 2^20 polynomials in total) as a strong-scaling record")
    ap.add_argument("--cfg5-polys", type=int, default=1 << 20)
    ap.add_argument("--cfg5-piece", type=int, default=16384)
    ap.add_argument("--cfg5-passes", type=int, default=5)
    ap.add_argument("--min-ms", type=float, default=60.0,
                    help="the timed region is extended (more steps) until it lasts at least this long")
    ap.add_argument("--tc-store", default="direct", choices=["direct", "staged"],
                    help="epilogue of the tensor-core kernel: 32 bytes per thread, or full-line stores through "
                         "shared memory (hbg_ctx_set_tc_store)")
    ap.add_argument("--sm-limit", type=int, default=0,
                    help="SMs the encode and the interpolation launches may use (0 = all): the rest stay free "
                         "for the gather's copy kernel (N > 1)")
    ap.add_argument("--sm-split", type=int, default=-1,
                    help="CTAs (SMs) of the encode launches; the interpolation gets the rest, so that the two "
                         "persistent kernels of a step run side by side (0 = both use every SM, back to back; "
                         "-1 = automatic at N = 1 and with the copy-engine gather: the split that balances the "
                         "two chains)")
    ap.add_argument("--graph-units", type=int, default=4,
                    help="steps per captured graph = lcm(sets, slots) x this (a replay ends with a join of "
                         "all streams, i.e. drains the gather pipeline once)")
    ap.add_argument("--no-graph", action="store_true", help="eager Python step loop instead of a CUDA graph")
    ap.add_argument("--matvec-path", default="auto",
                    help="auto (tensor-core kernel) | no-tc (the IMAD kernels of round 1) | ...")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
