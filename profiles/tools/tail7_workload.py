#!/usr/bin/env python
"""ncu workload: the perft tail at scale -- perft(7) of the opening three times (4.87 M ply-5 nodes, 38 000 tiles)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import leafline_b200 as ll

eng = ll.Engine(0)
for _ in range(3):
    assert eng.perft(ll.new_world(), 7)[0] == 3195901860
print("ok")
