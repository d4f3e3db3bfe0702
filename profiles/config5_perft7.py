#!/usr/bin/env python
"""BASELINE.json configs[4]: perft depth 7 from a tactics-heavy position (Kiwipete: ascension, secret service
and passing-by all live) across the GPUs of one box, 2-ply prefixes dealt round-robin, NCCL u64 all-reduce of
the per-root leaf counts.  Run under torchrun (or plain python for one GPU).

The faithful O(moves^2) oracle cannot reach this depth, so the count is verified hierarchically:
  (1) the all-reduced total equals the single-GPU count of the same tree (rank 0 recounts it alone),
  (2) it equals the sum over root moves of perft(6) of each child (a different split of the same tree),
  (3) a seeded sample of 2-ply prefixes is recounted at sub-depth 4 by the CPU oracle,
  (4) depths 1-3 equal the oracle's full counts (48 / 2 038 / 97 814 -- NOT orthodox chess: SURVEY.md 8(a) quirk 1).
"""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import leafline_b200 as ll
from leafline_b200 import sharding

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
DEPTH = int(os.environ.get("CONFIG5_DEPTH", "7"))

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
eng = ll.Engine(local)
stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)
eng.set_stream(stream.cuda_stream)
root = ll.reconstruct(KIWIPETE)
counts = torch.zeros(sharding.ROOT_SLOTS, dtype=torch.int64, device="cuda")


def sharded():
    leaves, per_root = eng.perft(root, DEPTH, shard=(rank, world))
    if world > 1:
        return sharding.reduce_leaf_counts(per_root, dist, torch, device="cuda", scratch=counts)
    return leaves, per_root


total, table = sharded()  # warm-up: sizes the level buffers
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
times = []
for _ in range(3):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record(stream)
    total, table = sharded()
    b.record(stream)
    b.synchronize()
    times.append(a.elapsed_time(b))
t = torch.tensor([sorted(times)[1]], dtype=torch.float64, device="cuda")
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
ms = float(t.item())

if rank == 0:
    import oracle as O

    out = {"config": f"perft({DEPTH}) Kiwipete, {world} GPU(s)", "fen": KIWIPETE, "depth": DEPTH, "n_gpus": world, "leaves": int(total),
           "ms": ms, "leaf_positions_per_sec": int(total) / (ms * 1e-3), "per_root": [int(x) for x in table[:48]]}
    t0 = time.perf_counter()
    alone, alone_roots = eng.perft(root, DEPTH)
    out["check_single_gpu_recount"] = bool(alone == total and list(alone_roots) == list(table[: len(alone_roots)]))
    out["single_gpu_seconds"] = time.perf_counter() - t0
    _, commits = eng.lookahead(root)
    by_child = [eng.perft(c["tree"], DEPTH - 1)[0] for c in commits]
    out["check_sum_of_children"] = bool(sum(by_child) == total and by_child == [int(x) for x in table[: len(by_child)]])
    rng = random.Random(7)
    prefixes = []
    for _ in range(6):
        c1 = commits[rng.randrange(len(commits))]
        _, seconds = eng.lookahead(c1["tree"])
        c2 = seconds[rng.randrange(len(seconds))]
        gpu = eng.perft(c2["tree"], 4)[0]
        cpu = int(O.perft_divide(c2["tree"], 4, threads=os.cpu_count() or 1).sum())
        prefixes.append({"prefix": ll.preserve(c2["tree"]), "gpu": gpu, "oracle": cpu})
    out["check_oracle_prefix_samples"] = all(p["gpu"] == p["oracle"] for p in prefixes)
    out["prefix_samples"] = prefixes
    shallow = [eng.perft(root, d)[0] for d in (1, 2, 3)]
    out["check_shallow_vs_oracle"] = shallow == [48, 2038, 97814] == [O.perft(O.reconstruct(KIWIPETE), d) for d in (1, 2, 3)]
    print(json.dumps(out))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
eng.close()
