#!/usr/bin/env python
"""Probe: build the library with -DLL_ORTHODOX_COP_CAPTURE -- the three secret-service rules in which the reference
differs from orthodox chess (b-file leer test, eligibility kept when a cop is stunned at home,
eligibility withdrawn by any cop leaving file a/h) switched to orthodox chess, nothing else changed -- and compare with
the published orthodox perft numbers; then the shipped build (the reference's rules)."""
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
print("opening   ", [eng.perft(ll.new_world(), d)[0] for d in (6, 7, 8)])
print("kiwipete  ", [eng.perft(ll.reconstruct("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"), d)[0] for d in (1, 2, 3, 4, 5, 6)])
print("position 3", [eng.perft(ll.reconstruct("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -"), d)[0] for d in (1, 2, 3, 4, 5, 6, 7)])
print("position 4", [eng.perft(ll.reconstruct("r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq -"), d)[0] for d in (1, 2, 3, 4, 5, 6)])
print("position 5", [eng.perft(ll.reconstruct("rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ -"), d)[0] for d in (1, 2, 3, 4, 5)])
print("position 6", [eng.perft(ll.reconstruct("r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - -"), d)[0] for d in (1, 2, 3, 4, 5, 6)])
''' % ROOT
for name, lib in (("orthodox cop capture (probe build)", VARIANT), ("shipped build (reference rules)", None)):
    env = dict(os.environ)
    if lib:
        env["LEAFLINE_B200_LIB"] = lib
    out = subprocess.run([sys.executable, "-c", CHILD], env=env, capture_output=True, text=True)
    print(name, "\n", out.stdout, out.stderr[-300:])
print("""published orthodox numbers (chessprogramming wiki, Perft Results):
opening    [119060324, 3195901860, 84998978956]
kiwipete   [48, 2039, 97862, 4085603, 193690690, 8031647685]
position 3 [14, 191, 2812, 43238, 674624, 11030083, 178633661]
position 4 [6, 264, 9467, 422333, 15833292, 706045033]
position 5 [44, 1486, 62379, 2103487, 89941194]
position 6 [46, 2079, 89890, 3894594, 164075551, 6923051137]""")
