#!/usr/bin/env python
"""ab8_ab.py <lib.so> ... -- the depth-8 search of the opening by the host alpha-beta driver (device_plies 4), five times per
build (seconds each; every build in its own process)"""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
CHILD = r'''
import sys, time
sys.path.insert(0, %r)
import leafline_b200 as ll
with ll.Engine(0) as eng:
    eng.alpha_beta_forecast(ll.new_world(), 5, device_plies=3)
    out = []
    for _ in range(5):
        t0 = time.perf_counter()
        fc, expanded, handed = eng.alpha_beta_forecast(ll.new_world(), 8, device_plies=4)
        out.append(time.perf_counter() - t0)
    eng.timing(True); eng.timing_reset()
    eng.alpha_beta_forecast(ll.new_world(), 8, device_plies=4)
    print(" ".join("%%.3f" %% s for s in out), "s; host nodes", expanded, "handed", handed, "score", float(fc[0]["score"]))
''' % ROOT
for lib in sys.argv[1:]:
    env = dict(os.environ, LEAFLINE_B200_LIB=os.path.abspath(lib))
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(f"{os.path.basename(lib):16s}", out.stdout.strip() or out.stderr.strip()[-600:], flush=True)
