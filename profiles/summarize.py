#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the text summaries committed here.

    python profiles/summarize.py launches gpurun_out/launches_r1.csv  > profiles/r1_launches_summary.txt
    python profiles/summarize.py kernel   gpurun_out/march_r1.ncu-rep > profiles/r1_march_ncu.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "sm__cycles_elapsed.avg.per_second", "launch__grid_size",
    "launch__block_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__shared_mem_per_block_dynamic",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__throughput.avg.pct_of_peak_sustained_active",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__pcsamp_warps_issue_stalled_math_pipe_throttle",
    "smsp__pcsamp_warps_issue_stalled_long_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_short_scoreboard",
    "smsp__pcsamp_warps_issue_stalled_wait",
    "smsp__pcsamp_warps_issue_stalled_barrier",
    "smsp__pcsamp_warps_issue_stalled_not_selected",
    "smsp__pcsamp_warps_issue_stalled_selected",
    "smsp__pcsamp_sample_count",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        agg.setdefault((row["Kernel Name"], row.get("Grid Size", "")), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({path})")
    print("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes")
    print(f"{'kernel':84s} {'grid':>16s} {'n':>5s} {'avg_us':>10s} {'share%':>7s}")
    for (k, g), v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:84]:84s} {g:>16s} {len(v):5d} {sum(v)/len(v)/1e3:10.2f} {sum(v)/tot*100:7.1f}")
    # the single-solve step of bench.py: one march launch + the back-transform launches that follow it
    step = collections.OrderedDict()
    for (k, g), v in agg.items():
        if "bldfm::" in k and g.endswith(", 1, 1)") and "k_march" in k:
            step[k] = sum(v) / len(v)
            n_single = len(v)
    for (k, g), v in agg.items():
        if "bldfm::" in k and "k_march" not in k and step and len(v) == n_single:
            step[k + " " + g] = sum(v) / len(v)
    if step:
        t = sum(step.values())
        print("# single-solve step (config 2): kernel shares")
        for k, v in step.items():
            print(f"#   {k[:90]:90s} {v/1e3:8.2f} us {v/t*100:6.1f} %")


def kernel(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full --clock-control none --import-source on  ({path})")
    for r in rows[2:]:
        print("==", r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:72s} {r[hdr.index(k)]:>18s} {units[hdr.index(k)]}")


if __name__ == "__main__":
    {"launches": launches, "kernel": kernel}[sys.argv[1]](sys.argv[2])
