#!/usr/bin/env python
"""Small pass over every kernel for compute-sanitizer (memcheck / racecheck): tiny sizes, all entry points."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import leafline_b200 as ll

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
eng = ll.Engine(0)
w = ll.reconstruct(KIWIPETE)
for depth in (1, 2, 3, 4, 4):
    eng.perft(w, depth)
eng.perft(w, 4, shard=(1, 3))
eng.perft(w, 4, shard=(1, 3))
out = torch.zeros(ll.PERFT_DEVICE_WORDS, dtype=torch.int64, device="cuda")
eng.perft_device(w, 4, out.data_ptr())
eng.synchronize()
soa = eng.playout_soa(300)
eng.synchronize()
worlds = eng.soa_download(soa)
vals = torch.empty(300, dtype=torch.float32, device="cuda")
eng.score_soa(soa, vals.data_ptr())
eng.horizon_soa(soa, 1, vals.data_ptr())
eng.soa_free(soa)
eng.score(worlds)
eng.lookahead(worlds[:100])
eng.reckless_lookahead(worlds[:100])
eng.in_critical_endangerment(worlds[:100], 0)
eng.horizon(worlds[:40], 2)
eng.horizon(worlds[:40], 1, quiet=2)
eng.kickoff(w, 3)
eng.kickoff(w, 2, quiet=2, shard=(0, 2))
eng.forecast(w, 3)
eng.close()
print("sanitizer workload ok")
