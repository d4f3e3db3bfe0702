"""world_size-2 (and 3) checks of the multi-rank host logic on the CPU box (gloo): unit ownership, the
u64 leaf-count all-reduce and the per-root score gather of leafline_b200/sharding.py.  The per-rank
compute here is the oracle standing in for the GPU (ll_perft / ll_kickoff apply the same unit rule on
the device; tests/test_gpu_parity.py checks that the device shards sum to the whole)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
POSITION_3 = "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -"
POSITION_4 = "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq -"


def free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def worker(rank, world, port, fen, depth, out_dir):
    import oracle as O
    from leafline_b200 import sharding

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    root = O.reconstruct(fen)
    roots = O.lookahead(root)
    # units = the nodes of ply perft_unit_ply(depth), each owned by (hash(parent) + index among the parent's commits) mod world
    # (the rule of ll_perft / ll_perft_multi; tests/test_gpu_parity.py checks the device against this restatement)
    unit_ply = sharding.perft_unit_ply(depth)
    assert unit_ply in (2, 3)
    per_root = np.zeros(len(roots), dtype=np.uint64)

    def descend(node, ply, root_index):
        for k, d in enumerate(O.lookahead(node)):
            if ply + 1 < unit_ply:
                descend(d["tree"], ply + 1, root_index)
            elif sharding.unit_owner(node, k, world) == rank:
                per_root[root_index] += np.uint64(O.perft(d["tree"], depth - unit_ply))

    for i, c in enumerate(roots):
        descend(c["tree"], 1, i)
    # make the sums cross 2^63 once to prove the u64-as-i64 transport is exact
    bias = np.uint64(0x7FFFFFFFFFFFFFFF) if rank == 0 else np.uint64(1)
    biased = per_root.copy()
    biased[0] += bias
    total, summed = sharding.reduce_leaf_counts(per_root, dist, torch)
    _, summed_biased = sharding.reduce_leaf_counts(biased, dist, torch)
    # root scores: each rank searched the roots it owns
    scores = np.full(len(roots), -np.inf, dtype=np.float32)
    for i in sharding.units_of(rank, world, len(roots)):
        scores[i] = -O.negamax(roots[i]["tree"], 1)
    gathered = sharding.gather_root_scores(scores, dist, torch)
    if rank == 0:
        np.save(os.path.join(out_dir, "summed.npy"), summed[: len(roots)])
        np.save(os.path.join(out_dir, "biased.npy"), summed_biased[: len(roots)])
        np.save(os.path.join(out_dir, "scores.npy"), gathered)
        np.save(os.path.join(out_dir, "total.npy"), np.array([total], dtype=np.uint64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,fen,depth", [(2, POSITION_4, 4), (3, POSITION_3, 5)])
def test_sharded_perft_and_kickoff_reduce_to_the_whole(world, fen, depth, tmp_path):
    import oracle as O

    O.build()
    mp.spawn(worker, args=(world, free_port(), fen, depth, str(tmp_path)), nprocs=world, join=True)
    root = O.reconstruct(fen)
    want = O.perft_divide(root, depth, threads=4)
    summed = np.load(tmp_path / "summed.npy")
    assert list(summed) == list(want) and int(np.load(tmp_path / "total.npy")[0]) == int(want.sum())
    biased = np.load(tmp_path / "biased.npy")
    expect0 = (int(want[0]) + 0x7FFFFFFFFFFFFFFF + (world - 1)) % (1 << 64)
    assert int(biased[0]) == expect0 and list(biased[1:]) == list(want[1:])
    roots, scores = O.kickoff(root, 2, False, threads=4)
    assert np.array_equal(np.load(tmp_path / "scores.npy"), scores)
