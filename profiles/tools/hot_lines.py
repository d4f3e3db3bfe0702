#!/usr/bin/env python
"""Join an ncu source-page CSV (per-SASS-instruction counters) with nvdisasm line info and print the
hottest source lines / opcodes of one kernel.

usage: hot_lines.py <report.ncu-rep> <lib.so> <kernel mangled-name substring> [launch index]
"""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

rep, lib, kname = sys.argv[1:4]
which = int(sys.argv[4]) if len(sys.argv) > 4 else 0
csv.field_size_limit(10**9)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
# split per kernel launch
launches, cur = [], None
for row in csv.reader(io.StringIO(src)):
    if row and row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        launches.append(cur)
    elif cur is not None and row:
        if cur["hdr"] is None:
            cur["hdr"] = row
        else:
            cur["rows"].append(row)
demangled = [l for l in launches if kname in l["name"]]
L = demangled[which]
hdr = L["hdr"]
ia, isrc, iinst, ithr = hdr.index("Address"), hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
isamp = hdr.index("# Samples")
base = int(L["rows"][0][ia], 16)
per_off = {}
for r in L["rows"]:
    per_off[int(r[ia], 16) - base] = (r[isrc].strip(), int(r[iinst] or 0), int(r[ithr] or 0), int(r[isamp] or 0))

with tempfile.TemporaryDirectory() as d:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=d, check=True, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout
# locate the function: mangled name guessed from template args in demangled name
sections = re.split(r"\n//-+ \.text\.(\S+) -+\n", dis)
funcs = dict(zip(sections[1::2], sections[2::2]))
cands = [k for k in funcs if all(tok in k for tok in os.environ.get("MANGLED", "").split(",") if tok)]
if len(cands) != 1:
    print("set MANGLED=<comma separated substrings> to pick one of:", *funcs.keys(), sep="\n  ")
    sys.exit(1)
line = ("?", 0)
by_line = collections.Counter()
by_line_samp = collections.Counter()
by_op = collections.Counter()
tot = 0
for ln in funcs[cands[0]].split("\n"):
    m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
    if m:
        line = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m:
        off = int(m.group(1), 16)
        if off in per_off:
            sass, inst, thr, samp = per_off[off]
            by_line[line] += inst
            by_line_samp[line] += samp
            op = sass.split()[0] if not sass.startswith("@") else sass.split()[1]
            by_op[op.split(".")[0]] += inst
            tot += inst
print(f"kernel {L['name']}: {tot} warp instructions")
print("-- by opcode")
for op, n in by_op.most_common(25):
    print(f"{n:12d} {100*n/tot:5.1f}%  {op}")
print("-- by source line (warp instructions, stall samples)")
srcs = {}
for (f, l), n in by_line.most_common(60):
    if f not in srcs:
        p = os.path.join(os.path.dirname(os.path.abspath(lib)), "csrc", f)
        srcs[f] = open(p).read().split("\n") if os.path.exists(p) else None
    text = srcs[f][l - 1].strip()[:110] if srcs[f] and l - 1 < len(srcs[f]) else ""
    print(f"{n:12d} {100*n/tot:5.1f}% samp {by_line_samp[(f,l)]:6d}  {f}:{l}  {text}")
