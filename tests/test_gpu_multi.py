"""Several GPUs sharing one tree behind the C ABI (ll_init_multi / ll_init_multi_rank, ll_perft_multi,
ll_perft_batch_multi, ll_kickoff_multi): results must equal the single-GPU entry points, which tests/test_gpu_parity.py
holds against the oracle.  On a one-GPU box the group's members share the device (the collective then runs over this
library's peer-memory reduction -- NCCL refuses two ranks on one device); with two or more GPUs the same checks run over
real peers, over NCCL and over peer memory, in one process and as one process per GPU."""
import os
import subprocess
import sys

import numpy as np
import pytest

import oracle as O
from positions import KIWIPETE, OPENING, POSITION_3, POSITION_4, POSITION_5

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def ll():
    import __graft_entry__ as entry

    entry.build()
    import leafline_b200

    return leafline_b200


@pytest.fixture(scope="module")
def eng(ll):
    e = ll.Engine(0)
    yield e
    e.close()


def gpu_count():
    import torch

    return torch.cuda.device_count()


def groups(ll):
    """device lists to test: members sharing GPU 0 always; distinct GPUs where the box has them"""
    out = [[0, 0], [0, 0, 0]]
    n = gpu_count()
    if n >= 2:
        out.append(list(range(min(n, 8))))
    return out


def test_unit_ownership_rule_matches_its_host_mirror(ll, eng):
    """Per-rank shard counts of the device against the oracle dealing the units by leafline_b200.sharding's restatement of
    the rule (hash(parent) + index among the parent's commits) mod world -- unit ply 2 and 3."""
    from leafline_b200 import sharding

    for fen, depth, world in ((POSITION_3, 4, 2), (POSITION_3, 5, 3), (POSITION_4, 4, 5)):
        root = O.reconstruct(fen)
        unit_ply = sharding.perft_unit_ply(depth)
        want = np.zeros(world, dtype=np.uint64)

        def descend(node, ply):
            for k, d in enumerate(O.lookahead(node)):
                if ply + 1 < unit_ply:
                    descend(d["tree"], ply + 1)
                else:
                    want[sharding.unit_owner(node, k, world)] += np.uint64(O.perft(d["tree"], depth - unit_ply))

        descend(root, 0)
        for synced in (True, False):
            eng.force_synced(synced)
            got = [eng.perft(ll.reconstruct(fen), depth, shard=(r, world))[0] for r in range(world)]
            assert got == [int(x) for x in want], (fen, depth, world, synced)
    eng.force_synced(False)


def test_perft_multi_equals_single_gpu(ll, eng):
    for devices in groups(ll):
        with ll.MultiEngine(devices) as g:
            assert g.world == len(devices) and g.local_count == len(devices)
            for fen, depth in ((OPENING, 6), (KIWIPETE, 5), (POSITION_5, 5), (OPENING, 5), (OPENING, 3), (OPENING, 1)):
                w = ll.reconstruct(fen)
                want = eng.perft(w, depth)
                for _ in range(3):  # first call sizes the levels (level-by-level), the next ones run the launch chain
                    got = g.perft(w, depth)
                    assert got[0] == want[0] and list(got[1]) == list(want[1]), (devices, fen, depth)
            g.set_shard_min_depth(perft_depth=2)  # deal even the shallow trees
            for depth in (2, 3, 4, 5):
                w = ll.reconstruct(KIWIPETE)
                want = eng.perft(w, depth)
                got = g.perft(w, depth)
                assert got[0] == want[0] and list(got[1]) == list(want[1]), (devices, depth)
            # an error on the way is voted on, not hung on
            bad = ll.new_world().copy()
            bad["plane"][1] |= bad["plane"][0]
            with pytest.raises(ll.LeaflineError):
                g.perft(bad, 6)
            assert g.perft(ll.new_world(), 5)[0] == 4865609


def test_perft_batch_multi(ll, eng):
    fens = [OPENING, KIWIPETE, POSITION_3, POSITION_4, POSITION_5, OPENING, KIWIPETE]
    worlds = np.array([ll.reconstruct(f) for f in fens], dtype=ll.WORLD_DTYPE)
    for devices in groups(ll):
        with ll.MultiEngine(devices) as g:
            for depth in (1, 3, 4):
                want = [eng.perft(w, depth)[0] for w in worlds]
                for _ in range(2):
                    assert [int(x) for x in g.perft_batch(worlds, depth)] == want, (devices, depth)
            weak = np.array([ll.new_world()] * len(devices), dtype=ll.WORLD_DTYPE)
            for _ in range(3):
                assert [int(x) for x in g.perft_batch(weak, 6)] == [119060324] * len(devices)


def test_kickoff_multi_equals_single_gpu(ll, eng):
    for devices in groups(ll):
        with ll.MultiEngine(devices) as g:
            g.set_shard_min_depth(kickoff_depth=2)
            for fen, depth, quiet in ((OPENING, 5, None), (KIWIPETE, 4, None), (KIWIPETE, 3, None), (KIWIPETE, 2, None), (POSITION_5, 4, 2),
                                      (POSITION_3, 6, None), ("7K/r7/1r6/8/8/8/8/7k b - -", 4, None), (OPENING, 1, None)):
                w = ll.reconstruct(fen)
                roots, scores, scored = eng.kickoff(w, depth, quiet=quiet)
                for _ in range(2):
                    r2, s2, n2 = g.kickoff(w, depth, quiet=quiet)
                    assert r2.tobytes() == roots.tobytes() and s2.tobytes() == scores.tobytes() and n2 == scored, (devices, fen, depth)
        with ll.MultiEngine(devices) as g:  # default threshold: a depth-5 tree is answered whole by every rank
            w = ll.new_world()
            roots, scores, scored = eng.kickoff(w, 5)
            r2, s2, n2 = g.kickoff(w, 5)
            assert s2.tobytes() == scores.tobytes() and n2 == scored


def test_group_of_one_rank(ll, eng):
    with ll.MultiEngine.for_rank(0, 0, 1, None) as g:
        assert g.world == 1
        assert g.perft(ll.new_world(), 5)[0] == 4865609
        assert g.kickoff(ll.new_world(), 4)[1].tobytes() == eng.kickoff(ll.new_world(), 4)[1].tobytes()


WORKER = r"""
import os, sys
sys.path.insert(0, REPO_ROOT)
import numpy as np, torch, torch.distributed as dist
import leafline_b200 as ll
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
uid = torch.zeros(128, dtype=torch.uint8, device="cuda")
if rank == 0:
    uid.copy_(torch.frombuffer(bytearray(ll.MultiEngine.unique_id()), dtype=torch.uint8))
dist.broadcast(uid, 0)
g = ll.MultiEngine.for_rank(rank, rank, world, bytes(uid.cpu().numpy()))
want_path = os.environ.get("LEAFLINE_B200_REDUCE")
if want_path == "nccl":
    assert g.reduce_path == "nccl"
e = ll.Engine(rank)
kiwi = ll.reconstruct("r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -")
for w, depth in ((ll.new_world(), 6), (kiwi, 5), (ll.new_world(), 7)):
    want = e.perft(w, depth)
    for _ in range(3):
        got = g.perft(w, depth)
        assert got[0] == want[0] and list(got[1]) == list(want[1]), (rank, depth)
weak = np.array([ll.new_world()] * world, dtype=ll.WORLD_DTYPE)
for _ in range(3):
    assert [int(x) for x in g.perft_batch(weak, 6)] == [119060324] * world
g.set_shard_min_depth(kickoff_depth=2)
for w, depth in ((ll.new_world(), 5), (kiwi, 4), (ll.new_world(), 6)):
    roots, scores, scored = e.kickoff(w, depth)
    for _ in range(2):
        r2, s2, n2 = g.kickoff(w, depth)
        assert r2.tobytes() == roots.tobytes() and s2.tobytes() == scores.tobytes() and n2 == scored, (rank, depth)
bad = ll.new_world().copy()
bad["plane"][1] |= bad["plane"][0]
try:
    g.perft(bad, 6)
    raise SystemExit("a bad root must fail on every rank")
except ll.LeaflineError:
    pass
assert g.perft(ll.new_world(), 5)[0] == 4865609
print("rank %d ok %s\n" % (rank, g.reduce_path), end="", flush=True)
g.close(); e.close()
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("reduce_path", ["nccl", "peer"])
def test_one_process_per_gpu(ll, reduce_path, tmp_path):
    """torchrun-style: every rank its own process and GPU, the NCCL id carried by torch.distributed, every call collective."""
    n = gpu_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = min(n, 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER.replace("REPO_ROOT", repr(ROOT)))
    env = dict(os.environ, LEAFLINE_B200_REDUCE=reduce_path)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", "29561", str(script)], env=env, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert out.stdout.count("ok") == world, out.stdout  # (the ranks' lines may interleave)
