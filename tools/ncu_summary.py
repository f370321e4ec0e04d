"""Summarise one or more `ncu --set full` captures into a small CSV: one column per kernel launch.
Accepts .ncu-rep reports (read through `ncu -i ... --page raw --csv`) or the .raw.csv files tools/ncu_export.sh
leaves behind.  ncu picks a unit PER VALUE (one launch reports ms, the next us; MB here, GB there), so every
duration is converted to microseconds and every byte count to bytes before it is written -- one unit per row.
Usage: python tools/ncu_summary.py out.csv capture1 [capture2 ...]"""
import csv
import subprocess
import sys

WANT = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "smsp__inst_executed.sum", "smsp__inst_executed_pipe_fp64.sum", "smsp__inst_executed_pipe_fma.sum",
    "smsp__thread_inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum",
]
TIME = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6, "nsecond": 1e-3}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def normalize(unit, value):
    """(unit, value) -> (unit, value) with durations in us and sizes in byte"""
    try:
        v = float(value.replace(",", ""))
    except ValueError:
        return unit, value
    if unit in TIME:
        return "us", "%.3f" % (v * TIME[unit])
    if unit in BYTES:
        return "byte", "%.0f" % (v * BYTES[unit])
    return unit, value


def read_columns(path):
    if path.endswith(".ncu-rep"):
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    else:
        txt = open(path).read()
    rows = list(csv.reader(txt.splitlines()))
    hdr, units = rows[0], rows[1]
    return [{h: normalize(u, v) if h != "Kernel Name" else (u, v) for h, u, v in zip(hdr, units, vals)} for vals in rows[2:]]


def write_summary(out, cols):
    keys = WANT + sorted(k for k in cols[0] if "issue_stalled" in k and k.endswith("per_issue_active.ratio"))
    with open(out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["metric", "unit"] + [c["Kernel Name"][1].split("(")[0] for c in cols])
        for k in keys:
            present = [c for c in cols if k in c]
            if present:
                w.writerow([k, present[0][k][0]] + [c.get(k, ("", ""))[1] for c in cols])


def main(out, paths):
    cols = []
    for p in paths:
        cols += read_columns(p)
    write_summary(out, cols)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2:])
