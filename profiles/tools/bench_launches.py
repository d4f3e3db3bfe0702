#!/usr/bin/env python
"""Format the ncu launch list of the bench command into profiles/<tag>_bench_launches.md.

usage: bench_launches.py <tag> <ncu csv of `bench.py --steps 2 --warmup 3 --no-extras`> [<bench json of the same build>]
"""
import csv
import json
import os
import re
import shutil
import sys

tag, path = sys.argv[1:3]
live = json.loads(open(sys.argv[3]).read().strip().splitlines()[-1]) if len(sys.argv) > 3 else None
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    if "at::" in name:
        name = "torch fill (L2 flush, untimed)"
    else:
        name = re.sub(r"\(.*", "", name.replace("void ", "").replace("ll::", "")).replace("(bool)", "").replace("(int)", "")
        name = name.replace("k_expand_records<1, 128>", "k_expand_records<PERFT>").replace("k_expand_records<0, 128>", "k_expand_records<EMIT>")
        m = re.match(r"k_expand<(\d), (\d), (\d+), (\d)>", name)
        if m:
            name = "k_expand<%s, %s>" % (["legal", "reckless"][int(m.group(1))], ["EMIT", "PERFT"][int(m.group(2))])
        name = name.replace("k_count<0>", "k_count<legal>").replace("k_count<1>", "k_count<reckless>")
    rows.append((name, r["Grid Size"], float(r["Metric Value"]) / 1e3))
# the last timed step: from the last single-block root launch to the end
start = max(i for i, r in enumerate(rows) if r[0] in ("k_store_world", "k_aos_to_soa"))
step = rows[start:]
total = sum(r[2] for r in step)
out = ["# %s: ncu launch list of `python bench.py --steps 2 --warmup 3 --no-extras`" % tag, "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none` (raw: `%s_bench_launches_ncu.csv`).  Per-launch times are cold-cache and" % tag,
       "serialised, so only the SHARES are comparable with the live CUDA-event timing of `bench.py`.", "",
       "One timed step (= one `ll_perft(opening, 6)` launch chain; live it is replayed as a CUDA graph), last step of the run:", "",
       "| launch | grid | us (ncu) | share |", "|---|---|---:|---:|"]
for name, grid, us in step:
    out.append("| `%s` | %s | %.1f | %.1f%% |" % (name, grid, us, 100 * us / total))
out.append("| total | | %.1f | |" % total)
out.append("")
if live:
    k = live.get("kernels") or {}
    tail = live["roofline"]["ms"]
    out.append("Live (`bench.py`, CUDA events on the launching stream, same build): step %.3f ms, perft tail %.3f ms (%.0f %%)%s."
               % (live["ms_per_step"], tail, 100 * tail / live["ms_per_step"],
                  ", narrow levels %.3f ms" % k["emit_legal"]["ms_per_step"] if "emit_legal" in k else ""))
    tail_ncu = sum(us for name, _, us in step if "PERFT" in name)
    out.append("Under ncu the tail is %.0f %% of the kernel time of the step: the shares agree." % (100 * tail_ncu / total))
    out.append("")
out += ["The first `ll_perft` call of a context takes the level-by-level path (it sizes the level buffers; `k_count` + `k_scan_block_sums` per ply);",
        "every later call is the chain above.", "", "## every launch in order (us)", "", "```"]
for i, (name, grid, us) in enumerate(rows):
    out.append("%4d %-40s grid %14s %10.1f" % (i, name, grid, us))
out.append("```")
open(os.path.join(ROOT, "profiles", "%s_bench_launches.md" % tag), "w").write("\n".join(out) + "\n")
shutil.copy(path, os.path.join(ROOT, "profiles", "%s_bench_launches_ncu.csv" % tag))
print("\n".join(out[:20]))
