"""Copies the evidence of one gpu_round.sh visit (gpurun_out/<tag>/) into profiles/ under a round/version label:
bench lines, launch list + shares, the ncu --set full summary of the dominant kernels (one unit per row, see
ncu_summary.py), emit_traffic.json (DRAM bytes per bin-update of the row kernel) and fused_fp64.json (FP64
instructions per bin-update of the fused kernel), which bench.py scales by the launch it times.
Usage: python tools/publish_profiles.py <tag> <label>      e.g.  r2n r02"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import ncu_summary  # noqa: E402

CAPTURES = ("prof_f64", "prof_f32", "prof_fused_f64", "prof_synth_f64", "prof_stream_call")
N_PROF, M_PROF = 262144, 4096


def main(tag, label):
    src = os.path.join(ROOT, "gpurun_out", tag)
    dst = os.path.join(ROOT, "profiles")
    for name, target in (("bench.json", "%s_bench.json"), ("bench_reference.json", "%s_bench_reference.json"),
                         ("launches_bench.csv", "%s_launches_bench.csv"), ("mid_sweep.md", "%s_mid_sweep.md"),
                         ("stream_sweep.md", "%s_stream_sweep.md")):
        if os.path.exists(os.path.join(src, name)):
            shutil.copy(os.path.join(src, name), os.path.join(dst, target % label))
    if os.path.exists(os.path.join(src, "launches_bench.csv")):
        shares = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "summarize_launches.py"),
                                 os.path.join(src, "launches_bench.csv")], capture_output=True, text=True).stdout
        open(os.path.join(dst, "%s_bench_kernel_shares.md" % label), "w").write(shares)
        print(shares)
    cols = []
    for f in CAPTURES:
        path = os.path.join(src, f + ".raw.csv")
        if os.path.exists(path) and os.path.getsize(path) > 0:
            cols += ncu_summary.read_columns(path)[:1]
    if not cols:
        return
    ncu_summary.write_summary(os.path.join(dst, "%s_ncu_full.csv" % label), cols)
    bin_updates = N_PROF * M_PROF
    for c in cols:
        name = c["Kernel Name"][1].split("(")[0]
        print(name[:70], c["gpu__time_duration.sum"], c["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"][1],
              c["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"][1], c["sm__warps_active.avg.pct_of_peak_sustained_active"][1])
    row = cols[0]
    rd, wr = float(row["dram__bytes_read.sum"][1]), float(row["dram__bytes_write.sum"][1])
    json.dump({"kernel": row["Kernel Name"][1].split("(")[0],
               "source": "profiles/%s_ncu_full.csv (ncu --set full, one launch, n=%d, m=%d)" % (label, N_PROF, M_PROF),
               "dram_bytes_read": rd, "dram_bytes_write": wr, "bin_updates": bin_updates,
               "dram_bytes_per_bin_update": (rd + wr) / bin_updates}, open(os.path.join(dst, "emit_traffic.json"), "w"), indent=1)
    fused = [c for c in cols if ", 0, 0, 2, " in c["Kernel Name"][1] or ", 0, false, 2, " in c["Kernel Name"][1]]
    if fused and "smsp__inst_executed_pipe_fp64.sum" in fused[0]:
        warp_instr = float(fused[0]["smsp__inst_executed_pipe_fp64.sum"][1])
        json.dump({"kernel": fused[0]["Kernel Name"][1].split("(")[0],
                   "source": "profiles/%s_ncu_full.csv (one launch of the fused analysis+synthesis kernel, n=%d, m=%d, hann, "
                             "latency 1)" % (label, N_PROF, M_PROF),
                   "smsp__inst_executed_pipe_fp64.sum": warp_instr, "bin_updates": bin_updates,
                   "fp64_thread_instructions_per_bin_update": warp_instr * 32.0 / bin_updates},
                  open(os.path.join(dst, "fused_fp64.json"), "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
