#!/usr/bin/env python
"""hz_ab.py <lib.so> ... -- horizon of 2^22 playout positions, one ply, per build (ms, median of 7; each in its own process)"""
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
soa = eng.playout_soa(1 << 22); vals = torch.empty(1 << 22, dtype=torch.float32, device="cuda")
eng.horizon_soa(soa, 1, vals.data_ptr())
eng.timing(True); eng.timing_reset()
for _ in range(7): scored = eng.horizon_soa(soa, 1, vals.data_ptr())
ms, n = eng.timing_query("horizon_last_ply")
print("k_horizon %%.3f ms  (%%.1f G leaf positions/s)  checksum %%r" %% (ms / n, scored / (ms / n) / 1e6, float(vals.sum())))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, LEAFLINE_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):16s}", out.stdout.strip() or out.stderr.strip()[-400:], flush=True)
