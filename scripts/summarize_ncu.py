#!/usr/bin/env python
"""Summarise ncu outputs into small text files under profiles/ (tracked).

  python scripts/summarize_ncu.py launches gpurun_out/launches.csv profiles/r01_launches.md
  python scripts/summarize_ncu.py rep gpurun_out/prof_gemm.ncu-rep profiles/r01_gemm_ncu.md
"""
import csv
import io
import re
import subprocess
import sys
from collections import defaultdict


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        txt = f.read()
    start = txt.find('"ID"')
    rd = csv.DictReader(io.StringIO(txt[start:]))
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        name = re.sub(r"\(.*", "", r["Kernel Name"])
        rows.append((name, ns))
    tot = sum(ns for _, ns in rows)
    agg = defaultdict(lambda: [0, 0.0])
    for n, ns in rows:
        agg[n][0] += 1
        agg[n][1] += ns
    with open(dst, "w") as f:
        f.write(f"# ncu launch list ({src}): {len(rows)} launches, {tot / 1e6:.2f} ms total (cold-cache, serialised: compare SHARES)\n\n")
        f.write("| kernel | launches | total ms | share | avg us |\n|---|---|---|---|---|\n")
        for n, (c, ns) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| {n} | {c} | {ns / 1e6:.3f} | {100 * ns / tot:.1f}% | {ns / c / 1e3:.1f} |\n")
    print(open(dst).read())


KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tensor_cycles_active", "sm__inst_executed_pipe_tensor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "lts__t_bytes.sum", "lts__t_sectors_op_read.sum",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__t_bytes.sum", "smsp__cycles_active.avg", "sm__cycles_elapsed.max",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic", "sm__pipe_tensor_subpipe",
        "smsp__warp_issue_stalled", "dram__throughput"]


def rep(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr_i = next(i for i, r in enumerate(rd) if r and r[0] == "ID")
    hdr, units, data = rd[hdr_i], rd[hdr_i + 1], rd[hdr_i + 2:]
    cols = [i for i, h in enumerate(hdr) if any(k in h for k in KEYS)]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary ({src})\n\n")
        for r in data:
            if len(r) < len(hdr):
                continue
            name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
            f.write(f"## {name[:100]}\n\n| metric | value | unit |\n|---|---|---|\n")
            for i in cols:
                f.write(f"| {hdr[i]} | {r[i]} | {units[i]} |\n")
            f.write("\n")
    print(open(dst).read()[:6000])


EXACT = ["gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
         "launch__shared_mem_per_block_dynamic", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum",
         "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
         "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
         "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
         "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
         "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
         "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
         "smsp__inst_executed.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
         "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio", "smsp__average_warp_latency_issue_stalled_short_scoreboard.ratio",
         "smsp__average_warp_latency_issue_stalled_barrier.ratio", "smsp__average_warp_latency_issue_stalled_lg_throttle.ratio",
         "smsp__average_warp_latency_issue_stalled_mio_throttle.ratio", "smsp__average_warp_latency_issue_stalled_math_pipe_throttle.ratio"]


def csvs(dst, *srcs):
    """raw-page CSVs exported on the GPU box (ncu -i X.ncu-rep --page raw --csv; the .ncu-rep files exceed the 64 MiB return cap)"""
    with open(dst, "w") as f:
        f.write("# ncu --set full --clock-control none, one launch per case (scripts/gpu_run.sh ncu-full; targets = scripts/prof_ops.py cases)\n\n")
        for src in srcs:
            rd = list(csv.reader(open(src, newline="")))
            hdr_i = next((i for i, r in enumerate(rd) if r and r[0] == "ID"), None)
            if hdr_i is None:
                f.write(f"## {src}: no data\n\n")
                continue
            hdr, units, data = rd[hdr_i], rd[hdr_i + 1], rd[hdr_i + 2:]
            for r in data:
                if len(r) < len(hdr):
                    continue
                name = r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?"
                case = re.sub(r".*ncu_[a-z0-9]+_|\.csv$", "", src)
                f.write(f"## {case}: `{name[:110]}`\n\n| metric | value | unit |\n|---|---|---|\n")
                for k in EXACT:
                    if k in hdr:
                        i = hdr.index(k)
                        f.write(f"| {k} | {r[i]} | {units[i]} |\n")
                f.write("\n")
    print(open(dst).read()[:5000])


if __name__ == "__main__":
    if sys.argv[1] == "csv":
        csvs(sys.argv[2], *sys.argv[3:])
    else:
        {"launches": launches, "rep": rep}[sys.argv[1]](sys.argv[2], sys.argv[3])
