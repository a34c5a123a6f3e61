"""Summarise `ncu --set full` reports into one CSV of the metrics the roofline discussion uses (one column per kernel launch).
Usage: python scripts/ncu_summary.py out.csv rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__lsu_writeback_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic"]


def main():
    out, reps = sys.argv[1], sys.argv[2:]
    cols, names = [], []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units, data = rows[0], rows[1], rows[2:]
        ix = {h: i for i, h in enumerate(hdr)}
        stall = [h for h in hdr if h.startswith("smsp__average_warps_issue_stalled_") and h.endswith("_per_issue_active.ratio")]
        for r in data:
            names.append(r[ix["Kernel Name"]].split("(")[0].replace("void <unnamed>::", ""))
            cols.append({h: (r[ix[h]], units[ix[h]]) for h in KEEP + stall if h in ix})
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + names)
        keys = [k for k in KEEP if any(k in c for c in cols)] + sorted({k for c in cols for k in c if k.startswith("smsp__average")})
        for k in keys:
            w.writerow([k, next((c[k][1] for c in cols if k in c), "")] + [c.get(k, ("", ""))[0] for c in cols])


if __name__ == "__main__":
    main()
