import sys, os, time
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import leafline_b200 as ll
import oracle as O
eng = ll.Engine(0)
n = 1 << 24
soa = eng.playout_soa(n)
pinned_in = torch.empty(n * 104, dtype=torch.uint8).pin_memory()
w = pinned_in.numpy().view(ll.WORLD_DTYPE)
for first in range(0, n, 1 << 20): w[first:first + (1 << 20)] = eng.soa_download(soa, first, 1 << 20)
out = torch.empty(n, dtype=torch.float32).pin_memory()
for rep in range(4):
    t0 = time.perf_counter(); eng.score_into(w, out.numpy()); dt = time.perf_counter() - t0
    print("e2e score 2^24: %.2f ms  %.1f M positions/s  %.1f GB/s over PCIe" % (dt * 1e3, n / dt / 1e6, n * 108 / dt / 1e9))
k = (1 << 20) * 3 + 12345   # ragged: 3 full chunks + a partial one
got = eng.score(w[:k])
assert got.tobytes() == O.score_batch(w[:k]).tobytes()
assert out.numpy()[:k].tobytes() == got.tobytes()
print("pipelined score ok")
