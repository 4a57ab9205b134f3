#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small tracked summaries under profiles/.

    python profiles/summarize.py launches gpurun_out/launches_X.csv profiles/rNN_launches_X.md
    python profiles/summarize.py full gpurun_out/prof_X.ncu-rep profiles/rNN_full_X.md

`launches`: per-kernel totals of gpu__time_duration.sum (cold-cache, serialised: compare SHARES).
`full`: the headline metrics of every profiled launch of an `ncu --set full` report (needs `ncu` on PATH).
"""
import collections
import csv
import io
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.sum.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.sum.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
]


def short(name):
    name = re.sub(r"\(.*", "", name)
    name = re.sub(r"<unnamed>::|void |mdseg::|\(anonymous namespace\)::", "", name)
    return name[:110]


def launches(src, dst):
    rows = [r for r in csv.reader(l for l in open(src) if l.startswith('"'))]
    hdr = rows[0]
    ik, iv, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot, cnt = collections.OrderedDict(), collections.Counter()
    for r in rows[1:]:
        v = float(r[iv].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[iu], 1.0)
        k = short(r[ik])
        tot[k] = tot.get(k, 0.0) + v
        cnt[k] += 1
    total = sum(tot.values())
    with open(dst, "w") as f:
        f.write(f"# ncu launch list `{src}` (gpu__time_duration.sum, --clock-control none)\n\n")
        f.write("Cold-cache, serialised replays: compare the SHARES, not the absolute times.\n\n")
        f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
        for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
            f.write(f"| `{k}` | {cnt[k]} | {v:.1f} | {100 * v / total:.1f}% |\n")
        f.write(f"\ntotal {total:.1f} us over {sum(cnt.values())} launches\n")


def full(src, dst):
    out = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    with open(dst, "w") as f:
        f.write(f"# ncu --set full summary of `{src}`\n")
        for r in rows[2:]:
            f.write(f"\n## `{short(r[hdr.index('Kernel Name')])}`\n\n| metric | value | unit |\n|---|---:|---|\n")
            for k in KEYS:
                if k in hdr:
                    i = hdr.index(k)
                    f.write(f"| {k} | {r[i]} | {units[i]} |\n")


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2], sys.argv[3])
