"""one-screen digest of a bench.py JSON line read from stdin"""
import json
import sys

for ln in sys.stdin:
    if not ln.startswith("{"):
        continue
    d = json.loads(ln)
    r = d.get("roofline", {})
    print(f"N={d['n_gpus']} value={d['value']:.3e} {d['unit']}  ms/step={d['ms_per_step']:.4f} steps={d['steps']} "
          f"loop={d['config'].get('step_loop')} par={d['config'].get('parallelism')}")
    print(f"   kernel_ms={r.get('kernel_ms')} frac={r.get('frac'):.3f} step_frac={r.get('step_frac', 0):.3f} "
          f"nvlink={r.get('nvlink', {}).get('frac')} floor_ms={r.get('nvlink', {}).get('floor_ms_per_step')}")
    print(f"   e2e={d['e2e']['value']:.3e} ({d['e2e']['ms_per_step']:.3f} ms)  clocks={d['clocks']}")
    c5 = d.get("cfg5_strong")
    if c5:
        print("   cfg5:", {k: c5[k] for k in c5 if k in ("error", "ms_per_pass", "value", "gather", "polys_per_gpu",
                                                          "kernel_ms_per_piece", "nvlink_floor_ms")})
