"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) into a markdown table of per-kernel shares.
Usage: python scripts/summarise_launches.py gpurun_out/launches.csv "title" [bench.json] > profiles/rNN_xxx_launches_summary.md"""
import csv
import json
import re
import sys
from collections import defaultdict


def main():
    path, title = sys.argv[1], sys.argv[2]
    rows = []
    with open(path, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") == "gpu__time_duration.sum":
            ns = float(r["Metric Value"]) * (1000.0 if r["Metric Unit"] in ("us", "usecond") else 1.0)
            rows.append((r["Kernel Name"], ns))
    # torch's own kernels (pool generation, fills, the TF32-peak matmul bench.py times for its GEMM fraction) are not the step
    own = [(n, ns) for n, ns in rows if not re.search(r"cutlass|at::|cublas|elementwise|distribution", n)]
    foreign = sum(ns for _, ns in rows) - sum(ns for _, ns in own)
    rows = own
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for name, ns in rows:
        short = re.sub(r"\(.*", "", name)
        short = short.replace("void ", "").replace("at::native::", "")
        agg[short][0] += 1
        agg[short][1] += ns
    print(f"# {title}\n")
    print("Command: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv python bench.py --steps 2 --warmup 3 "
          "--no-cpu-baseline`  \n(cold-cache, serialised per-launch times: compare SHARES, not absolutes; first 400 launches, "
          "including torch's setup fills)\n")
    print(f"Kernels of torch itself (synthetic pool, fills, the TF32-peak matmul of the bench) are left out: {foreign / 1000:.0f} us.\n")
    print("| kernel | launches | total us | share |\n|---|---|---|---|")
    for name, (n, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        if ns / tot < 0.001:
            continue
        print(f"| `{name[:90]}` | {n} | {ns / 1000:.1f} | {100 * ns / tot:.1f}% |")
    if len(sys.argv) > 3:
        b = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1])
        st = b["roofline"]["stages_ms"]
        print(f"\nbench.py (same build, not under ncu): {b['ms_per_step']:.3f} ms/step, {b['value'] / 1e6:.1f} M frames/s, "
              f"e2e {b['e2e']['value'] / 1e6:.1f} M frames/s; event-timed stages (ms): "
              + ", ".join(f"{k} {v:.3f}" for k, v in st.items()))


if __name__ == "__main__":
    main()
