"""The driver-facing contract of bench.py that can be checked without a GPU: the
reference arm prints ONE JSON line with the agreed keys (it times the CPU
restatement of the reference's NTL path), non-zero ranks of a multi-rank
reference run stay silent, and the B200 arm refuses to run without a device."""

import json
import os
import subprocess
import sys

import torch
from conftest import ROOT

BENCH = os.path.join(ROOT, "bench.py")


def run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, BENCH] + args, capture_output=True, text=True, cwd=ROOT,
                          env=e, timeout=600)


def test_reference_arm_line():
    res = run(["--impl", "reference", "--gpus", "1", "--steps", "2", "--warmup", "1"])
    assert res.returncode == 0, res.stderr[-2000:]
    lines = [ln for ln in res.stdout.splitlines() if ln.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["scaling"] == "weak"
    assert d["metric"] == "GF(p) share reconstructions/sec at n=16,t=5" and d["unit"] == "shares/s"
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["steps"] >= 1 and d["n_gpus"] == 1
    assert d["vs_baseline"] is None and d["data"] == "synthetic"
    assert "workload" in d["config"]
    cb = d["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == d["value"] and cb["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": 0}


def test_reference_arm_ignores_torchrun_thread_cap():
    """torchrun exports OMP_NUM_THREADS=1; the CPU baseline must still use every core of the box
    (round 1's N >= 2 reference lines ran on one thread and inflated the ratio ~16x)"""
    res = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--batch", "4096"],
              env={"RANK": "0", "WORLD_SIZE": "2", "LOCAL_RANK": "0", "OMP_NUM_THREADS": "1"})
    assert res.returncode == 0, res.stderr[-2000:]
    d = json.loads([ln for ln in res.stdout.splitlines() if ln.startswith("{")][0])
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert d["config"]["batch_polys_per_step"] == 4096


def test_reference_arm_other_ranks_are_silent():
    res = run(["--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1"],
              env={"RANK": "1", "WORLD_SIZE": "2", "LOCAL_RANK": "1"})
    assert res.returncode == 0 and res.stdout.strip() == ""


def test_b200_arm_needs_a_device():
    if torch.cuda.is_available():
        return
    res = run(["--steps", "1", "--warmup", "1", "--no-cpu"])
    assert res.returncode != 0
    assert "no CPU fallback" in res.stderr


def test_sm_split_choice():
    """the automatic --sm-split: balanced chains, the documented 104 + 44 at the headline size"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.choose_sm_split(512, 148) == 104
    for tiles, n_sm in ((512, 148), (8192, 148), (100, 148), (512, 132), (1, 148)):
        e = bench.choose_sm_split(tiles, n_sm)
        assert n_sm // 2 <= e < n_sm - 8
        # no neighbouring split has a shorter longer-chain
        cost = lambda x: max(-(-tiles // x) * 2.4, -(-tiles // (n_sm - x)) * 1.0)  # noqa: E731
        assert all(cost(e) <= cost(x) for x in range(n_sm // 2, n_sm - 8))


def test_numa_helpers():
    """sysfs parsing of the e2e leg's NUMA pinning; unknown topology is a no-op that restores affinity"""
    import importlib.util

    spec = importlib.util.spec_from_file_location("bench_mod2", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    assert bench.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert bench.parse_cpulist("5") == {5} and bench.parse_cpulist("") == set()

    class NoSuchGpu:
        pci_domain_id, pci_bus_id, pci_device_id = 0xFFFF, 0xFE, 0x1F

    assert bench.gpu_numa_cpus(NoSuchGpu) == (None, None)
    before = os.sched_getaffinity(0)
    pin = bench.numa_local(NoSuchGpu)
    with pin:
        assert os.sched_getaffinity(0) == before
    assert os.sched_getaffinity(0) == before and pin.describe()["applied"] is False
