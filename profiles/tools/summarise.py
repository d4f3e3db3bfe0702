#!/usr/bin/env python
"""Turn ncu captures (gpurun_out/) into the committed summaries under profiles/.

usage: summarise.py <round tag, e.g. r02> <launches.csv> <hot.ncu-rep> [<pipes.csv>]
writes profiles/<tag>_launches.md, profiles/<tag>_kernels.csv, profiles/constants.json
(<pipes.csv>: `ncu --metrics sm__inst_executed_pipe_{alu,fma,lsu,xu}.sum,... --csv` of the same workload: warp instructions per pipe)
"""
import collections
import csv
import io
import json
import os
import re
import subprocess
import sys

tag, launches_csv, rep = sys.argv[1:4]
pipes_csv = sys.argv[4] if len(sys.argv) > 4 else None
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
OUT = os.path.join(ROOT, "profiles")
csv.field_size_limit(10**9)


def short(name):
    name = name.replace("void ", "").replace("ll::", "")
    name = re.sub(r"\(.*", "", name)
    name = name.replace("(bool)", "").replace("(int)", "")
    m = re.match(r"k_expand<(\d), (\d)(?:, (\d+), (\d))?>", name)
    if m:
        name = "k_expand<%s, %s>" % (["legal", "reckless"][int(m.group(1))], ["EMIT", "PERFT"][int(m.group(2))])
        if m.group(3) and m.group(3) != "128":
            name = name[:-1] + ", P=%s>" % m.group(3)
    m = re.match(r"k_expand_records<(\d), (\d+)>", name)
    if m:
        name = "k_expand_records<%s%s>" % (["EMIT", "PERFT"][int(m.group(1))], "" if m.group(2) == "128" else ", P=" + m.group(2))
    m = re.match(r"k_count<(\d)>", name)
    if m:
        name = "k_count<%s>" % ["legal", "reckless"][int(m.group(1))]
    return name


# ---- launch list
rows = [r for r in csv.reader(open(launches_csv)) if len(r) > 10 and r[0].isdigit()]
agg = collections.OrderedDict()
for r in rows:
    k = short(r[4])
    agg.setdefault(k, [0, 0.0])
    agg[k][0] += 1
    agg[k][1] += float(r[-1]) / 1e3
tot = sum(v[1] for k, v in agg.items() if k != "k_playout")  # the playout generator builds the workload; it is not on the hot path
with open(os.path.join(OUT, f"{tag}_launches.md"), "w") as f:
    f.write(f"# {tag}: ncu launch list of `python profiles/workload.py`\n\n")
    f.write("`ncu --metrics gpu__time_duration.sum --clock-control none` -- per-launch times are cold-cache and serialised: "
            "compare SHARES, not absolutes.  Workload: 3x perft(6) of the opening (1st call level-by-level, then the launch chain), "
            "2x horizon of 2^20 playout positions, 2x score of 2^24 playout positions (+ the playout generators).\n\n")
    f.write("| kernel | launches | total us | share |\n|---|---:|---:|---:|\n")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        share = "(workload generator)" if k == "k_playout" else f"{100 * v[1] / tot:.1f}%"
        f.write(f"| `{k}` | {v[0]} | {v[1]:.1f} | {share} |\n")
    f.write("\n## every launch in order (us)\n\n```\n")
    for r in rows:
        f.write(f"{int(r[0]):4d} {short(r[4]):45s} grid {r[8]:>14s} block {r[7]:>12s} {float(r[-1]) / 1e3:10.1f}\n")
    f.write("```\n")

# ---- full-set metrics of the hot kernels
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr, units = rr[0], rr[1]
want = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__inst_executed.sum", "smsp__thread_inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__warps_eligible.avg.per_cycle_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum",
]
idx = [hdr.index(w) for w in want if w in hdr]
UNIT = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0}
consts = {}
with open(os.path.join(OUT, f"{tag}_kernels.csv"), "w", newline="") as f:
    w = csv.writer(f)
    w.writerow(["launch"] + [hdr[i] for i in idx])
    w.writerow(["unit"] + [units[i] for i in idx])
    for n, r in enumerate(rr[2:]):
        w.writerow([n] + [r[i] for i in idx])
        name = short(r[hdr.index("Kernel Name")])

        def val(metric):
            i = hdr.index(metric)
            return float(r[i]) * UNIT.get(units[i], 1.0)

        dram = val("dram__bytes_read.sum") + val("dram__bytes_write.sum")
        tinst = float(r[hdr.index("smsp__inst_executed.sum")]) * float(r[hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio")])
        key = {"k_expand<legal, PERFT>": "perft6_tail", "k_expand_records<PERFT>": "perft6_tail", "k_horizon<0>": "horizon_2p20", "k_horizon<false>": "horizon_2p20", "k_horizon<0, 0>": "horizon_2p20", "k_horizon<false, false>": "horizon_2p20", "k_score": "score"}.get(name)
        if key:  # the last launch of each kind wins (steady state)
            consts[f"{key}_dram_bytes"] = dram
            consts[f"{key}_thread_inst"] = tinst
            consts[f"{key}_warp_inst"] = float(r[hdr.index("smsp__inst_executed.sum")])
            consts[f"{key}_ncu_seconds"] = val("gpu__time_duration.sum")
            consts[f"{key}_alu_pipe_pct"] = float(r[hdr.index("sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active")])
            consts[f"{key}_issue_active_pct"] = float(r[hdr.index("smsp__issue_active.avg.pct_of_peak_sustained_active")])
            consts[f"{key}_lanes_active_of_32"] = float(r[hdr.index("smsp__thread_inst_executed_per_inst_executed.ratio")])
if pipes_csv:  # warp instructions per pipe, the last launch of each kind
    with open(pipes_csv, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        name = short(r["Kernel Name"])
        key = {"k_expand_records<PERFT>": "perft6_tail", "k_horizon<0, 0>": "horizon_2p20", "k_horizon<false, false>": "horizon_2p20", "k_score": "score"}.get(name)
        m = re.match(r"sm__inst_executed_pipe_(\w+)\.sum", r["Metric Name"])
        if key and m:
            consts[f"{key}_{m.group(1)}_pipe_warp_inst"] = float(r["Metric Value"].replace(",", ""))
import hashlib

h = hashlib.sha256()
for name in ("ll_device.cuh", "ll_kernels.cuh"):
    with open(os.path.join(ROOT, "leafline_b200", "csrc", name), "rb") as f:
        h.update(f.read())
consts["device_code_sha256"] = h.hexdigest()  # bench.py withholds the instruction-count rooflines when the sources no longer hash to this
consts["source"] = f"profiles/{tag}_kernels.csv (ncu --set full of profiles/workload.py)"
json.dump(consts, open(os.path.join(OUT, "constants.json"), "w"), indent=1)
print(json.dumps(consts, indent=1))
