#!/usr/bin/env python
"""SASS evidence per hot kernel of libhbmpc_b200.so: instruction counts that prove the Blackwell
paths (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG / UBLKCP = TMA, SYNCS = mbarrier), the
integer-pipe work (IMAD.WIDE) and the absence of spills (STL / LDL).

    python tools/sass_excerpt.py > profiles/r2_sass_excerpt.md
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "honeybadgermpc_b200", "libhbmpc_b200.so")
WANT = ["UTCIMMA", "UTCHMMA", "LDTM", "UTMALDG", "UBLKCP", "UTCATOMSWS", "SYNCS", "LDGSTS", "IMAD.WIDE",
        "IMAD", "IADD3", "STG.E.ENL2.256", "STG", "LDG", "STL", "LDL", "MULTIMEM", "REDG", "ST.E", "MEMBAR",
        "BAR.SYNC", "NANOSLEEP"]
KERNELS = ["tc_apply_kernel<hb::FieldBLS", "interp_small_kernel<hb::FieldBLS, 6, 64, 2, false, 0>",
           "ntt16_g4_kernel<hb::FieldBLS, 6, 64, false>", "apply_matrix_smem_kernel<hb::FieldBLS>",
           "ntt_smem_kernel<hb::FieldBLS>", "gather_copy_signal_kernel", "gather_bulk_signal_kernel", "gather_copy_kernel",
           "gather_signal_arrived_kernel", "gather_wait_released_kernel", "gather_wait_kernel",
           "gather_release_kernel", "compare_columns_kernel", "columns_to_rows_kernel",
           "fnt_scale_scatter_kernel<hb::FieldBLS>", "fnt_pointwise_kernel<hb::FieldBLS>",
           "gao_kernel<hb::FieldBLS>", "wb_kernel<hb::FieldBLS>"]

def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, cur = {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = name
            counts[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur:
            op = m.group(1)
            counts[cur]["total"] += 1
            for w in WANT:
                if op == w or op.startswith(w + "."):
                    counts[cur][w] += 1
    print("# SASS excerpt of honeybadgermpc_b200/libhbmpc_b200.so (round 2)\n")
    print(f"`cuobjdump -sass`: architectures in the fat binary: {', '.join(arch)}; "
          f"{len(counts)} kernels.  Counts are static instructions per kernel instance "
          "(prefix match: `IMAD` includes `IMAD.WIDE`, `STG` includes `STG.E.ENL2.256`).\n")
    cols = [w for w in WANT if any(c[w] for c in counts.values())]
    print("| kernel instance | total | " + " | ".join(cols) + " |")
    print("|---|---:|" + "---:|" * len(cols))
    for k in KERNELS:
        for name in sorted(counts):
            if k in name:
                short = re.sub(r"hb::|\(.*\)$|void ", "", name)
                short = short.replace("FieldBLS", "BLS")[:70]
                c = counts[name]
                print(f"| `{short}` | {c['total']} | " + " | ".join(str(c[w]) if c[w] else "" for w in cols) + " |")
    print("\nReading: `tc_apply_kernel` carries the tcgen05 path (UTCIMMA = `tcgen05.mma.kind::i8`, LDTM = "
          "`tcgen05.ld`, UTMALDG = `cp.async.bulk.tensor`, UBLKCP = `cp.async.bulk`, UTCATOMSWS = TMEM "
          "allocation, SYNCS = mbarrier operations) and no local-memory spills (no STL / LDL).")


if __name__ == "__main__":
    sys.exit(main())
