#!/usr/bin/env python
"""ncu workload: what ONE rank of an eight-GPU group runs for perft(7) of the opening (its share of the ply-3 units),
three times -- the launch list names where a strong-scaling step's time goes (everything but the reduction kernel)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import leafline_b200 as ll

eng = ll.Engine(0)
total = 0
for _ in range(3):
    total = eng.perft(ll.new_world(), 7, shard=(3, 8))[0]
print("rank 3 of 8 counts", total, "of 3195901860 leaves")
