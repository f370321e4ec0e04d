"""Summarise one or more `ncu --set full` reports (.ncu-rep) into a small CSV: one column per kernel launch.
Usage: python tools/ncu_summary.py out.csv rep1.ncu-rep [rep2.ncu-rep ...]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum",
]


def main(out, reps):
    cols = []
    for rep in reps:
        txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        hdr, units = rows[0], rows[1]
        for vals in rows[2:]:
            d = {h: (u, v) for h, u, v in zip(hdr, units, vals)}
            cols.append(d)
    keys = WANT + sorted(k for k in cols[0] if "issue_stalled" in k and k.endswith("per_issue_active.ratio"))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [c["Kernel Name"][1].split("(")[0] for c in cols])
        for k in keys:
            if k in cols[0]:
                w.writerow([k, cols[0][k][0]] + [c.get(k, ("", ""))[1] for c in cols])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
