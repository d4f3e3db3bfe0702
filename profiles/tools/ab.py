#!/usr/bin/env python
"""A/B timing of library builds: ab.py <lib.so> [<lib.so> ...]  (each runs in its own process).
Reports perft(6)/perft(7) of the opening (ms, L2 flushed between runs) and the horizon of 2^22 playouts."""
import os
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import statistics, sys, time
sys.path.insert(0, %r)
import torch
import leafline_b200 as ll
eng = ll.Engine(0)
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream); eng.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
root = ll.new_world()
def timed(fn, reps):
    ts = []
    for _ in range(reps):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(stream); fn(); b.record(stream); b.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), min(ts)
for _ in range(5): assert eng.perft(root, 6)[0] == 119060324
p6 = timed(lambda: eng.perft(root, 6), 40)
for _ in range(2): assert eng.perft(root, 7)[0] == 3195901860
p7 = timed(lambda: eng.perft(root, 7), 7)
eng.timing(True); eng.timing_reset()
for _ in range(5): eng.perft(root, 6)
tail = eng.timing_query("perft_tail")[0] / 5; emit = eng.timing_query("emit_legal")[0] / 5
eng.timing(False)
soa = eng.playout_soa(1 << 22); vals = torch.empty(1 << 22, dtype=torch.float32, device="cuda")
eng.horizon_soa(soa, 1, vals.data_ptr())
hz = timed(lambda: eng.horizon_soa(soa, 1, vals.data_ptr()), 5)
print("perft6 %%.4f (min %%.4f)  tail %%.4f emit %%.4f | perft7 %%.3f | horizon %%.3f" %% (p6[0], p6[1], tail, emit, p7[0], hz[0]))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, LEAFLINE_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):14s}", out.stdout.strip() or out.stderr.strip()[-400:], flush=True)
