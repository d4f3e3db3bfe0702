#!/usr/bin/env python
"""ncu workload for the horizon kernel alone: 2^20 playout positions, one ply (LEAFLINE_B200_LIB picks the build)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import leafline_b200 as ll

eng = ll.Engine(0)
soa = eng.playout_soa(1 << 20)
vals = torch.empty(1 << 20, dtype=torch.float32, device="cuda")
for _ in range(2):
    scored = eng.horizon_soa(soa, 1, vals.data_ptr())
eng.synchronize()
print("scored", scored)
