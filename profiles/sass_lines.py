#!/usr/bin/env python
"""Attribute the per-SASS-instruction counters of an ncu report to CUDA source lines.

    python profiles/sass_lines.py <report.ncu-rep> <cubin> <kernel-substring> [top_n]

ncu's `--page source --csv` lists SASS rows in program order; `nvdisasm -g` lists the same
instructions with `//## File "...", line N` markers.  The two are zipped by position.
"""
import collections
import csv
import io
import re
import subprocess
import sys

rep, cubin, kern = sys.argv[1:4]
top_n = int(sys.argv[4]) if len(sys.argv) > 4 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + (sys.argv[5] if len(sys.argv) > 5 else kern)],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr, data = rows[1], rows[2:]
iex, ismp, isrc = hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("Source")
dis = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.splitlines()
# locate the kernel's text section
lines, on, cur = [], False, ("?", 0)
for l in dis:
    if l.startswith(".text.") or ".section" in l and ".text." in l:
        on = kern in l
        continue
    if not on:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append((cur, l.split("*/", 1)[1].strip().rstrip(";").strip()))
if len(lines) != len(data):
    print(f"warning: {len(lines)} disassembled instructions vs {len(data)} profiled rows", file=sys.stderr)
agg = collections.defaultdict(lambda: [0, 0])
n = min(len(lines), len(data))
tot = sum(int(r[iex]) for r in data)
tots = sum(int(r[ismp]) for r in data)
for (loc, txt), r in zip(lines[:n], data[:n]):
    agg[loc][0] += int(r[iex])
    agg[loc][1] += int(r[ismp])
print(f"total warp instructions {tot}, samples {tots}")
for loc, (e, s) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top_n]:
    print(f"{loc[0]:18s} L{loc[1]:<5d} {100 * e / tot:5.1f}% inst {100 * s / max(tots, 1):5.1f}% samples")
