#!/usr/bin/env python
"""Turns the raw ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

    python profiles/summarize.py launches gpurun_out/launches.csv            > profiles/<tag>_launches.txt
    python profiles/summarize.py kernel   gpurun_out/prof_trace.ncu-rep      > profiles/<tag>_k_trace_ncu.txt
    python profiles/summarize.py traffic  gpurun_out/prof_trace.ncu-rep WORKLOAD N_GPUS [OUT.json]  (updates profiles/k_trace_traffic.json)
    python profiles/summarize.py sass     single-file-vulkan-pathtracing_b200/lib/libbpt.so > profiles/k_trace.sass

`traffic` averages, over ALL captured k_trace launches (capture every bounce of a frame: `-k regex:k_trace -s <first
launch of a timed frame> -c <launches per frame>`), the DRAM bytes (dram__bytes_read.sum + dram__bytes_write.sum), the
duration and the issue / ALU-pipe / L2-hit figures, and stores them with the sha256 of csrc/ the capture was made
from: bench.py prints `roofline.traffic` only when that hash is the hash of the code it runs.
"""
import collections
import csv
import io
import json
import os
import re
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
    """path: a .ncu-rep, or the `ncu -i rep --page raw --csv` export of one (reports of many launches are too large to
    bring back from the box; their CSV export is not)."""
    raw = open(path).read() if path.endswith(".csv") else \
        subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print(f"# ncu --set full --clock-control none: {path} ({len(data)} launches)")
    name = hdr.index("Kernel Name")
    print("kernel:", data[0][name][:100])
    for k in KEYS:
        if k in hdr:
            i = hdr.index(k)
            print(f"{k:72s} {units[i]:16s} " + "  ".join(d[i] for d in data))


def _num(x):
    try:
        return float(x.replace(",", ""))
    except ValueError:
        return float("nan")


def traffic(path, workload, n_gpus, out=None):
    here = os.path.dirname(os.path.abspath(__file__))
    sys.path.insert(0, os.path.dirname(here))
    import bench
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = lambda k: [_num(d[hdr.index(k)]) for d in data]
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    def nbytes(k):
        u = units[hdr.index(k)]
        return [v * scale.get(u, 1.0) for v in col(k)]
    tscale = {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0, "nsecond": 1e-9, "usecond": 1e-6, "msecond": 1e-3, "second": 1.0}
    dur = [v * tscale.get(units[hdr.index("gpu__time_duration.sum")], 1e-9) for v in col("gpu__time_duration.sum")]
    dram = [a + b for a, b in zip(nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum"))]
    wavg = lambda k: sum(v * t for v, t in zip(col(k), dur)) / sum(dur) / 100.0 if k in hdr else None
    entry = {"workload": workload, "n_gpus": int(n_gpus), "csrc_hash": bench.csrc_hash(), "source": os.path.basename(path),
             "kernel": data[0][hdr.index("Kernel Name")][:80], "launches": len(data),
             "dram_bytes_per_launch": sum(dram) / len(dram), "dram_bytes_total": sum(dram),
             "duration_ms_per_launch_under_ncu": 1e3 * sum(dur) / len(dur),
             "dram_gbs_under_ncu": sum(dram) / sum(dur) / 1e9,
             "issue_active_frac": wavg("smsp__issue_active.avg.pct_of_peak_sustained_active"),
             "alu_pipe_frac": wavg("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active"),
             "fma_pipe_frac": wavg("sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active"),
             "l2_hit_rate": wavg("lts__t_sector_hit_rate.pct"),
             "dram_pct_of_peak_ncu": wavg("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
             "per_launch": [{"ms": 1e3 * t, "dram_bytes": b} for t, b in zip(dur, dram)]}
    out = out or os.path.join(here, "k_trace_traffic.json")
    try:
        entries = json.load(open(out))
    except (OSError, ValueError):
        entries = []
    entries = [e for e in entries if (e["workload"], e["n_gpus"]) != (workload, int(n_gpus))] + [entry]
    json.dump(entries, open(out, "w"), indent=1)
    print(json.dumps({k: v for k, v in entry.items() if k != "per_launch"}, indent=1))


def sass(lib):
    """SASS of the traversal kernel instance the big scenes run (global records, one level, no counters)."""
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    blocks = re.split(r"(?=\t\tFunction : )", txt)
    for b in blocks:
        m = re.match(r"\t\tFunction : (\S+)", b)
        # STAGED, TWO_LEVEL, COUNT, FUSED all false: the instance bpt_trace runs on the big scenes
        if m and "k_trace" in m.group(1) and "ELi8ELb0ELb0ELb0ELb0E" in m.group(1):
            print(b.rstrip())
            return
    raise SystemExit("k_trace<.., false, false, false, false> not found")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel, "traffic": traffic, "sass": sass}[sys.argv[1]](*sys.argv[2:])
