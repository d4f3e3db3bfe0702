#!/usr/bin/env python
"""ncu workload for the perft tail alone: perft(6) of the opening three times (LEAFLINE_B200_LIB picks the build)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import leafline_b200 as ll

eng = ll.Engine(0)
for _ in range(3):
    assert eng.perft(ll.new_world(), 6)[0] == 119060324
print("ok")
