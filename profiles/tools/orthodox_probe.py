#!/usr/bin/env python
"""Probe: build the library with -DLL_ORTHODOX_COP_CAPTURE (eligibility withdrawn when a cop's home square is captured
on, as in orthodox chess -- the ONE rule of the reference that can change opening perft counts up to ply 8) and compare
with the published orthodox perft numbers; then the shipped build, whose surplus is the reference's phantom-cop rule."""
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
VARIANT = "/tmp/libleafline_orthodox.so"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-Xcompiler", "-fPIC", "-shared",
                "-DLL_ORTHODOX_COP_CAPTURE", "-o", VARIANT, os.path.join(ROOT, "leafline_b200", "csrc", "ll_api.cu")], check=True)
CHILD = r'''
import sys
sys.path.insert(0, %r)
import leafline_b200 as ll
eng = ll.Engine(0)
print("opening", [eng.perft(ll.new_world(), d)[0] for d in (6, 7, 8)])
print("position 5", [eng.perft(ll.reconstruct("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ -"), d)[0] for d in (1, 2, 3, 4)])
''' % ROOT
for name, lib in (("orthodox cop capture (probe build)", VARIANT), ("shipped build (reference rules)", None)):
    env = dict(os.environ)
    if lib:
        env["LEAFLINE_B200_LIB"] = lib
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(name, "\n", out.stdout, out.stderr[-300:])
print("published orthodox: opening 119060324 / 3195901860 / 84998978956; position 5: 44 / 1486 / 62379 / 2103487")
