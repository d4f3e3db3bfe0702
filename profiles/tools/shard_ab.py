#!/usr/bin/env python
"""shard_ab.py <lib.so> ... -- one rank's share of perft(7) of the opening dealt over 8 (and 4, 2) ranks, ms per build"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import statistics, sys
sys.path.insert(0, %r)
import torch
import leafline_b200 as ll
eng = ll.Engine(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
root = ll.new_world()
out = []
for world in (8, 4, 2, 1):
    for _ in range(3): eng.perft(root, 7, shard=(world // 2, world))
    ts = []
    for _ in range(9):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); eng.perft(root, 7, shard=(world // 2, world)); b.record(stream); b.synchronize()
        ts.append(a.elapsed_time(b))
    out.append("1/%%d: %%.3f ms" %% (world, statistics.median(ts)))
print("  ".join(out))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, LEAFLINE_B200_LIB=os.path.abspath(lib))
    o = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):18s}", o.stdout.strip() or o.stderr.strip()[-400:], flush=True)
