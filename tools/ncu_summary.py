#!/usr/bin/env python
"""Summarise ncu outputs brought back in gpurun_out/ into small tracked files under profiles/.

    python tools/ncu_summary.py launches gpurun_out/<tag>_launches.csv profiles/<tag>_launches.md
    python tools/ncu_summary.py full     gpurun_out/<tag>_full_<kernel>.ncu-rep profiles/<tag>_full_<kernel>.md

`launches`: per-kernel totals / shares of the `--metrics gpu__time_duration.sum` pass (cold-cache, serialised: shares
are meaningful, absolutes are not).  `full`: the counters named in /opt/skills/guides/B200_PROFILING.md for one
`--set full` capture.
"""
import csv
import re
import subprocess
import sys
from collections import OrderedDict


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = name.replace("void ", "").replace("pmgt::", "")
    return name[:90]


def launches(src, dst):
    rows = []
    with open(src, newline="") as f:
        lines = [ln for ln in f if ln.startswith('"')]
    rd = csv.DictReader(lines)
    for r in rd:
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r["Metric Unit"]
        v_us = v / 1e3 if unit == "ns" else (v if unit in ("us", "usecond") else v * 1e3)
        rows.append((short(r["Kernel Name"]), v_us, r["Grid Size"], r["Block Size"]))
    agg = OrderedDict()
    for n, us, g, b in rows:
        d = agg.setdefault(n, [0, 0.0, g, b])
        d[0] += 1
        d[1] += us
    tot = sum(d[1] for d in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list summary of `{src}`\n\n")
        f.write(f"{len(rows)} launches, {tot / 1e3:.3f} ms of kernel time (cold-cache, serialised under ncu: compare shares)\n\n")
        f.write("| kernel | launches | total us | avg us | share | grid | block |\n|---|---:|---:|---:|---:|---|---|\n")
        for n, d in sorted(agg.items(), key=lambda kv: -kv[1][1]):
            f.write(f"| `{n}` | {d[0]} | {d[1]:.1f} | {d[1] / d[0]:.2f} | {100 * d[1] / tot:.1f}% | {d[2]} | {d[3]} |\n")
    print(open(dst).read()[:3000])


KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram__cycles_active.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__occupancy_limit_registers",
    "launch__occupancy_limit_shared_mem", "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "sm__cycles_elapsed.max", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
]


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    lines = [ln for ln in out.splitlines() if ln.startswith('"')]
    rd = list(csv.reader(lines))
    hdr, units, data = rd[0], rd[1], rd[2:]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}`\n\n")
        for row in data:
            name = row[hdr.index("Kernel Name")]
            f.write(f"## `{short(name)}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {row[i]} | {units[i]} |\n")
            try:
                rd_b = float(row[hdr.index("dram__bytes_read.sum")].replace(",", ""))
                wr_b = float(row[hdr.index("dram__bytes_write.sum")].replace(",", ""))
                u = units[hdr.index("dram__bytes_read.sum")]
                f.write(f"\ntraffic (read + write) = {rd_b + wr_b:.3f} {u}\n\n")
            except Exception:
                pass
    print(open(dst).read()[:4000])


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
