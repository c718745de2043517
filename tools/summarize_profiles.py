#!/usr/bin/env python
"""Turn raw ncu outputs in gpurun_out/ into the small text summaries committed under profiles/.

    python tools/summarize_profiles.py launches gpurun_out/launches_stage1.csv profiles/r1_launches_stage1.txt
    python tools/summarize_profiles.py full gpurun_out/med_full.ncu-rep profiles/r1_med_full.txt
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.sum.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__cycles_elapsed.avg.per_second", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_tc_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_tc.sum", "smsp__cycles_active.avg"]


def launches(src, dst):
    lines = [l for l in open(src) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        v = v / 1e3 if row["Metric Unit"] == "ns" else (v * 1e3 if row["Metric Unit"] == "ms" else v)
        name = re.sub(r"\(.*", "", row["Kernel Name"])[:110]
        agg[name][0] += 1
        agg[name][1] += v
        tot += v
    ours = sum(t for k, (n, t) in agg.items() if "faln::" in k or "m3::" in k)
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --metrics gpu__time_duration.sum --clock-control none; cold-cache, serialised)\n")
        f.write(f"# total {tot:.1f} us over {sum(n for n, _ in agg.values())} launches; libfalnet kernels {ours:.1f} us ({100 * ours / tot:.1f}%)\n")
        f.write("#   time_us  launches  share  kernel\n")
        for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"{t:10.1f} {n:6d} {100 * t / tot:6.2f}%  {k}\n")


def full(src, dst):
    """src: a .ncu-rep, or the CSV of its raw page (`ncu -i X.ncu-rep --page raw --csv`, exported on the GPU box when the
    report itself is too large to bring back)."""
    if src.endswith(".csv"):
        raw = open(src).read()
    else:
        raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# source: {src} (ncu --set full --clock-control none --import-source on), one block per profiled launch\n")
        for r in rows[2:]:
            d = dict(zip(hdr, r))
            f.write(f"\n== {d.get('Kernel Name', '?')}  grid {d.get('Grid Size', '')} block {d.get('Block Size', '')}\n")
            for k in KEYS:
                if k in d:
                    f.write(f"{k:80s} {d[k]:>18s} {units[hdr.index(k)]}\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
