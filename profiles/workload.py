#!/usr/bin/env python
"""Small fixed workload for ncu captures: the hot kernels once each, in steady state.
perft(6) of the opening (k_expand<legal,EMIT> x4 + k_expand<legal,PERFT>), horizon of 2^20 playout
positions (k_expand<reckless,HORIZON>), score of 2^24 playout positions (k_score)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import leafline_b200 as ll

eng = ll.Engine(0)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
root = ll.new_world()
for _ in range(3):  # 1st call sizes the levels (level-by-level path), 2nd/3rd are the launch chain
    leaves, _ = eng.perft(root, 6)
    assert leaves == 119060324
soa = eng.playout_soa(1 << 20)
vals = torch.empty(1 << 20, dtype=torch.float32, device="cuda")
for _ in range(2):
    eng.horizon_soa(soa, 1, vals.data_ptr())
eng.soa_free(soa)
soa = eng.playout_soa(1 << 24)
out = torch.empty(1 << 24, dtype=torch.float32, device="cuda")
for _ in range(2):
    eng.score_soa(soa, out.data_ptr())
eng.synchronize()
eng.soa_free(soa)
eng.close()
print("workload ok")
