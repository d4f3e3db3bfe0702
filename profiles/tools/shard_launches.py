#!/usr/bin/env python
"""Format the ncu launch list of profiles/tools/shard_workload.py into profiles/<tag>_shard_launches.md.
usage: shard_launches.py <tag> <ncu csv>"""
import csv
import os
import re
import sys

tag, path = sys.argv[1:3]
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if l.startswith('"')]
for r in csv.DictReader(lines):
    if r["Metric Name"] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"].replace("void ", "").replace("ll::", ""))
    name = name.replace("k_expand_records<1, 128>", "k_expand_records<PERFT>").replace("k_expand_records<0, 128>", "k_expand_records<EMIT>")
    rows.append((name, r["Grid Size"], float(r["Metric Value"]) / 1e3))
start = max(i for i, r in enumerate(rows) if r[0] == "k_store_world")  # the last call
step = rows[start:]
tot = sum(r[2] for r in step)
what = {0: "root handed over as a kernel argument", 1: "ply 0 -> 1 (20 nodes), root moves rank-sorted", 2: "ply 1 -> 2 (400 nodes), every rank expands all",
        3: "ply 2 -> 3: every rank expands all 400 parents, KEEPS its own ~1 110 of the 8 902 units (the shard rule fused into the expansion)",
        4: "ply 3 -> 4 (this rank's ~25 k nodes)", 5: "ply 4 -> 5 (this rank's ~610 k nodes)", 6: "the tail: ply-5 nodes expanded on chip, ply-7 leaves bulk-counted"}
out = ["# %s: one rank's share of a strong-scaling step under ncu" % tag, "",
       "`ncu --metrics gpu__time_duration.sum --clock-control none python profiles/tools/shard_workload.py` -- `ll_perft(opening, 7, shard 3 of 8)` on ONE GPU:",
       "exactly the launches rank 3 of an eight-GPU group issues for `ll_perft_multi(opening, 7)`, minus the packing and the reduction kernel (`k_peer_allreduce`,",
       "one block, ~3 us + the wait for the slowest rank).  Cold-cache, serialised times: shares, not absolutes.  Rank 3 counts 405 634 028 of the",
       "3 195 901 860 leaves (12.69 %; an even eighth is 12.5 %; the eight ranks hold 12.34-12.69 %, `tools/shares.py`).", "",
       "| launch | grid | us (ncu) | share | what |", "|---|---|---:|---:|---|"]
for i, (n, g, us) in enumerate(step):
    out.append("| `%s` | %s | %.1f | %.1f%% | %s |" % (n, g, us, 100 * us / tot, what.get(i, "")))
out.append("| total | | %.1f | | |" % tot)
out += ["", "The replicated head (plies 0-2 and the dealing) is %.1f us of %.1f; the rest scales with 1/N.  Live (`bench.py`, CUDA events, `profiles/r02_scale_n8.json`):" % (sum(r[2] for r in step[:4]), tot),
        "the whole step takes 0.915 ms on eight GPUs against 5.81 ms on one.", ""]
open(os.path.join(ROOT, "profiles", "%s_shard_launches.md" % tag), "w").write("\n".join(out) + "\n")
print("\n".join(out))
