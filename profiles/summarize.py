#!/usr/bin/env python
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv            > profiles/<tag>_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_trace.ncu-rep      > profiles/<tag>_k_trace_ncu.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "launch__registers_per_thread", "launch__block_size", "launch__grid_size", "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
    "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for x in csv.DictReader(lines):
        k = x["Kernel Name"].split("(")[0].replace("void ", "").replace("<unnamed>::", "")
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += float(x["Metric Value"].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none (cold-cache, serialised): {path}")
    print(f"{'kernel':44s} {'launches':>8s} {'total ms':>10s} {'avg us':>10s} {'share':>7s}")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k:44s} {a[0]:8d} {a[1] / 1e6:10.3f} {a[1] / a[0] / 1e3:10.1f} {a[1] / tot:7.3f}")


def kernel(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full --clock-control none: {path} ({len(data)} launches)")
    name = hdr.index("Kernel Name")
    print("kernel:", data[0][name][:100])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:72s} {units[i]:16s} " + "  ".join(d[i] for d in data))


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
