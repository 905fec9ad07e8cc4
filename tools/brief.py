"""print the key numbers of a bench.py JSON line read from stdin"""
import json
import sys

tag = sys.argv[1] if len(sys.argv) > 1 else ""
for line in sys.stdin:
    line = line.strip()
    if not line.startswith("{"):
        if line:
            print(tag, "|", line[:300])
        continue
    d = json.loads(line)
    r = d.get("roofline") or {}
    e = d.get("e2e") or {}
    print(tag, "value=%.3e" % d["value"], "ms/step=%.4f" % d["ms_per_step"],
          "host_ms/step=%.4f" % d.get("host_enqueue_ms_per_step", 0), "kernels=", r.get("kernel_ms"),
          "frac=%.3f" % r.get("frac", 0), "e2e_ms=%.3f" % e.get("ms_per_step", 0), "n_gpus=", d.get("n_gpus"))
