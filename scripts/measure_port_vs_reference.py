#!/usr/bin/env python3
"""CPU container only: time the UNMODIFIED reference's train step (imported from /root/reference with the four shims) and
the numpy port (oracle/st_oracle.py) on the same batch and the same host cores, and record the ratio.  bench.py's reference
arm runs the port on the GPU box (the reference tree does not travel) and prints this ratio next to its number.

    python scripts/measure_port_vs_reference.py [B=200] [steps=3]   ->  profiles/r02_cpu_port_vs_reference.json"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from oracle import ref_loader  # noqa: E402
from signaltrain_b200 import data  # noqa: E402


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 200
    steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
    threads = os.cpu_count() or 1
    st = ref_loader.load_reference()
    if st is None:
        raise SystemExit("reference tree not found")
    wl = bench.WORKLOADS[1]
    from oracle import st_oracle as O
    d = O.model_dims(wl["scale"], bench.SHRINK, wl["knobs"])
    x, y, k = data.make_pool(B, d.C, d.L, getattr(data, wl["effect"])(), bench.SR, seed=218)
    ref_fps, ref_ms = ref_loader.reference_step_rate(st, x, y, k, steps, 1, threads)
    port_fps, port_ms = bench.cpu_reference_step_rate(B, steps, 1, threads, wl)
    out = {"batch": B, "steps": steps, "threads": threads, "reference_ms_per_step": ref_ms, "port_ms_per_step": port_ms,
           "port_over_reference_time": port_ms / ref_ms, "where": "build container (no GPU)",
           "note": "reference = unmodified /root/reference signaltrain (torch CPU, all threads); port = oracle/st_oracle.py Trainer (numpy float32)"}
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    with open(os.path.join(ROOT, "profiles", "r02_cpu_port_vs_reference.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
