"""Copies the evidence of one gpu_round.sh visit (gpurun_out/<tag>/) into profiles/ under a version label:
bench lines, launch list + shares, the ncu --set full summary of the four dominant kernels, emit_traffic.json.
Usage: python tools/publish_profiles.py <tag> <label>      e.g.  v18 v4"""
import csv
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__inst_executed.sum", "sm__cycles_elapsed.avg.per_second", "lts__t_bytes.sum"]


def main(tag, label):
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    shutil.copy(os.path.join(src, "bench.json"), os.path.join(dst, "r01_bench_%s.json" % label))
    shutil.copy(os.path.join(src, "bench_reference.json"), os.path.join(dst, "r01_bench_reference_%s.json" % label))
    shutil.copy(os.path.join(src, "launches_bench.csv"), os.path.join(dst, "r01_launches_bench_%s.csv" % label))
    shares = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"),
                             os.path.join(src, "launches_bench.csv")], capture_output=True, text=True).stdout
    open(os.path.join(dst, "r01_bench_%s_kernel_shares.md" % label), "w").write(shares)
    cols = []
    for f in ("prof_f64", "prof_f32", "prof_fused_f64", "prof_synth_f64"):
        rows = list(csv.reader(open(os.path.join(src, f + ".raw.csv"))))
        cols.append(dict(zip(rows[0], zip(rows[1], rows[2]))))
    keys = WANT + sorted(k for k in cols[0] if "issue_stalled" in k and k.endswith("per_issue_active.ratio"))
    with open(os.path.join(dst, "r01_ncu_full_%s_n262144_m4096_hann.csv" % label), "w", newline="") as fh:
        w = csv.writer(fh)
        w.writerow(["metric", "unit"] + [c["Kernel Name"][1].split("(")[0] for c in cols])
        for k in keys:
            if k in cols[0]:
                w.writerow([k, cols[0][k][0]] + [c.get(k, ("", ""))[1] for c in cols])
    c = cols[0]
    n, m = 262144, 4096
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = float(c["dram__bytes_read.sum"][1]) * unit[c["dram__bytes_read.sum"][0]]
    wr = float(c["dram__bytes_write.sum"][1]) * unit[c["dram__bytes_write.sum"][0]]
    json.dump({"kernel": c["Kernel Name"][1].split("(")[0],
               "source": "profiles/r01_ncu_full_%s_n262144_m4096_hann.csv (ncu --set full, one launch, n=262144, m=4096)" % label,
               "dram_bytes_read": rd, "dram_bytes_write": wr, "bin_updates": n * m,
               "dram_bytes_per_bin_update": (rd + wr) / (n * m)}, open(os.path.join(dst, "emit_traffic.json"), "w"), indent=1)
    print(shares)
    for col in cols:
        print(col["Kernel Name"][1][:60], col["gpu__time_duration.sum"], col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][1],
              col["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][1])


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
