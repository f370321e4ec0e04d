"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches, time, share."""
import csv
import sys

OURS = ("phase_table_kernel", "delta_kernel", "chunk_totals_kernel", "carry_scan_kernel", "emit_kernel",
        "synth_kernel", "scan_emit_kernel", "roundtrip_kernel", "convolve_kernel")


def main(path):
    rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 5]
    hdr, agg = None, {}
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr is None:
            continue
        d = dict(zip(hdr, r))
        name = d["Kernel Name"].split("(")[0].replace("void ", "").replace("sdftb200::", "")
        agg.setdefault(name, []).append(float(d["Metric Value"].replace(",", "")))
    ours = {k: v for k, v in agg.items() if any(k.startswith(o) for o in OURS)}
    tot = sum(sum(v) for v in ours.values()) or 1.0
    print("| kernel | launches | total ms | avg us | share of sdft_b200 kernel time |")
    print("|---|---:|---:|---:|---:|")
    for k, v in sorted(ours.items(), key=lambda kv: -sum(kv[1])):
        print("| %s | %d | %.3f | %.1f | %.1f%% |" % (k, len(v), sum(v) / 1e6, sum(v) / len(v) / 1e3, 100 * sum(v) / tot))
    other = sum(sum(v) for k, v in agg.items() if k not in ours)
    print("\nother kernels (torch fill/copy etc.): %.3f ms over %d launches" % (other / 1e6, sum(len(v) for k, v in agg.items() if k not in ours)))


if __name__ == "__main__":
    main(sys.argv[1])
