"""Multi-GPU plumbing of the leaf layer: which rank owns which unit, and the two tiny collectives.

The path shards by independent subtrees (the reference itself runs one thread per root move,
mind.rs:399-414): no position ever crosses NVLink.  The only exchanges are a u64 sum of per-root leaf
counts and a max-gather of per-root f32 scores, both <= 8 KiB and latency-bound.  `dist` is
torch.distributed (NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np

ROOT_SLOTS = 1024  # ll_perft / ll_kickoff root capacity


_M64 = (1 << 64) - 1


def unit_hash(parent):
    """`unit_hash` of csrc/ll_device.cuh restated on a WORLD_DTYPE record (host-side mirror, used by the CPU tests)."""
    meta = int(parent["initiative"]) | (int(parent["service_eligibility"]) << 1) | (int(parent["passing_by"]) << 8)
    h = ((meta & 0xFFFF) * 0x9E3779B97F4A7C15) & _M64
    for plane in parent["plane"]:
        h = (((h ^ int(plane)) * 0xD6E8FEB86659FD93) + (h >> 29)) & _M64
    return h >> 32


def perft_unit_ply(depth):
    """The ply whose nodes are the work units of a sharded ll_perft (0: the tree is too shallow to deal, rank 0 counts it)."""
    tail_ply = depth - 2 if depth >= 3 else depth - 1
    s = min(3, tail_ply)
    return s if s >= 2 else 0


def unit_owner(parent, c, world_size):
    """Rank that owns the unit made by the c-th commit (canonical order) of `parent` -- the rule ll_perft(shard_rank,
    shard_world) and ll_perft_multi apply on the device: a pure function of the tree, nothing is exchanged."""
    return (unit_hash(parent) + c) % world_size


def units_of(rank, world_size, n_units):
    """Units dealt by index (the root-move sharding of ll_kickoff; the canonical levels of ll_kickoff_multi)."""
    return range(rank, n_units, world_size)


def reduce_leaf_counts(per_root, dist, torch, device="cpu", scratch=None):
    """Sum per-root u64 leaf counts over all ranks -> (total, per_root[ROOT_SLOTS] as uint64).
    u64 sums travel as two's-complement i64 (NCCL/gloo have no unsigned 64-bit sum)."""
    host = torch.zeros(ROOT_SLOTS, dtype=torch.int64)
    a = np.ascontiguousarray(per_root, dtype=np.uint64)
    host[: len(a)] = torch.from_numpy(a.view(np.int64))
    buf = host.to(device) if scratch is None else scratch.copy_(host, non_blocking=True)
    dist.all_reduce(buf)
    out = buf.cpu().numpy().view(np.uint64)
    return int(out.sum(dtype=np.uint64)), out


def gather_root_scores(scores, dist, torch, device="cpu"):
    """Each rank holds -inf for the root moves it did not search (ll_kickoff): element-wise max over ranks."""
    buf = torch.full((ROOT_SLOTS,), float("-inf"), dtype=torch.float32)
    buf[: len(scores)] = torch.from_numpy(np.ascontiguousarray(scores, dtype=np.float32))
    buf = buf.to(device)
    dist.all_reduce(buf, op=dist.ReduceOp.MAX)
    return buf.cpu().numpy()[: len(scores)]
