"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle.
Integer / byte / index results must be bit-exact; f32 scores are compared bit-for-bit too (the
north-star tolerance is 1e-5 relative; the kernel reproduces the reference's rounding sequence)."""
import numpy as np
import pytest

import oracle as O
from positions import (KIWIPETE, NAMED, OPENING, POSITION_3, POSITION_4, POSITION_5, POSITION_6, named_worlds,
                       reckless_walk_worlds)

pytestmark = pytest.mark.gpu

SCORE_RTOL = 1e-5  # BASELINE.json north_star


@pytest.fixture(scope="module")
def eng():
    import __graft_entry__ as entry

    entry.build()
    import leafline_b200 as ll

    e = ll.Engine(0)
    yield e
    e.close()


@pytest.fixture(scope="module")
def playouts():
    return O.playout_batch(0x1EAF11E, 0, 1024)


@pytest.fixture(scope="module")
def reckless_worlds():
    return reckless_walk_worlds(seed=5, games=30, plies=50)


def assert_scores(got, want):
    assert got.dtype == np.float32
    np.testing.assert_allclose(got, want, rtol=SCORE_RTOL, atol=0)
    assert got.tobytes() == want.tobytes(), "scores are within tolerance but not bit-exact"


# ---------------------------------------------------------------- generated tables (furniture/tablemaker.py:41-62)
def test_device_tables_equal_the_reference_generator(eng):
    """The movement tables resident on the device, entry by entry, against the output of the reference's own generator
    (tests/golden/reference_tables.json, written by tests/golden/make_tables.py from furniture/tablemaker.py)."""
    import json
    import os

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_tables.json")) as f:
        ref = json.load(f)
    pony, fig = eng.debug_tables()
    assert [int(x) for x in pony] == ref["pony_movement_table"]
    assert [int(x) for x in fig] == ref["figurehead_movement_table"]


# ---------------------------------------------------------------- score (mind.rs:50-143)
def test_score_golden_bit_patterns(eng):
    import leafline_b200 as ll

    cases = {
        OPENING: 0x3F000000,
        "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3": 0xBECCCCCD,
        "3q1rk1/2R1bppp/pP2p3/N2b4/1r6/4BP2/1P1Q2PP/R5K1 b - -": 0xBEFFCCCC,
        KIWIPETE: 0x3DCCCCC7,
        "8/q1P1k/8/8/8/8/6PP/7K w - -": 0xC0600CCD,
        "k7/pp6/8/8/8/P7/P7/K7 w - -": 0x00000000,
    }
    worlds = np.array([ll.reconstruct(f) for f in cases], dtype=ll.WORLD_DTYPE)
    got = eng.score(worlds)
    assert [int(x) for x in got.view(np.uint32)] == list(cases.values())


def test_score_matches_oracle_on_all_position_sets(eng, playouts, reckless_worlds):
    for worlds in (named_worlds(), playouts, reckless_worlds, playouts[:1], playouts[:3]):
        assert_scores(eng.score(worlds), O.score_batch(worlds))


def test_score_empty_batch(eng):
    assert len(eng.score(np.zeros(0, dtype=O.WORLD_DTYPE))) == 0


def test_score_resident_soa_matches_host_path(eng, playouts):
    import torch

    soa = eng.soa_upload(playouts[:777])
    out = torch.empty(777, dtype=torch.float32, device="cuda")
    eng.score_soa(soa, out.data_ptr())
    eng.synchronize()
    assert_scores(out.cpu().numpy(), O.score_batch(playouts[:777]))
    back = eng.soa_download(soa)
    assert back.tobytes() == playouts[:777].tobytes()
    eng.soa_free(soa)


# ---------------------------------------------------------------- lookahead (life.rs:1052-1084)
def check_lookahead(eng, worlds, nihilistically):
    offsets, commits = eng.lookahead(worlds, nihilistically=nihilistically)
    assert offsets[0] == 0 and len(offsets) == len(worlds) + 1
    ref = O.reckless_lookahead if nihilistically else O.lookahead
    for i, w in enumerate(worlds):
        want = ref(w)
        got = commits[offsets[i]:offsets[i + 1]]
        assert len(got) == len(want), (i, O.preserve(w), len(got), len(want))
        assert got.tobytes() == want.tobytes(), (i, O.preserve(w))
    counts = eng.lookahead_counts(worlds, nihilistically=nihilistically)
    assert list(counts) == list(np.diff(offsets.astype(np.int64)))


@pytest.mark.parametrize("nihilistically", [False, True])
def test_lookahead_named_positions(eng, nihilistically):
    check_lookahead(eng, named_worlds(), nihilistically)


@pytest.mark.parametrize("nihilistically", [False, True])
def test_lookahead_playout_positions(eng, playouts, nihilistically):
    check_lookahead(eng, playouts[:400], nihilistically)


@pytest.mark.parametrize("nihilistically", [False, True])
def test_lookahead_reckless_walk_positions(eng, reckless_worlds, nihilistically):
    # no-figurehead boards, phantom cops, resurrected figureheads
    check_lookahead(eng, reckless_worlds[:500], nihilistically)


def test_lookahead_known_answers_of_the_reference(eng):
    import leafline_b200 as ll

    # life.rs:1407-1432: 20 opening moves, pony order b1a3 b1c3 g1f3 g1h3
    offsets, commits = eng.lookahead(ll.new_world())
    assert offsets[1] == 20
    ponies = [(ll.to_algebraic(c["whence"]), ll.to_algebraic(c["whither"])) for c in commits if c["job"] == ll.PONY]
    assert ponies == [("b1", "a3"), ("b1", "c3"), ("g1", "f3"), ("g1", "h3")]
    # life.rs:1360-1387: ascension order Pony, Scholar, Cop, Princess
    w = np.zeros((), dtype=ll.WORLD_DTYPE)
    w["plane"][0] = 1 << 48
    w["passing_by"] = 0xFF
    _, commits = eng.reckless_lookahead(w)
    assert [c["ascension"] & 7 for c in commits] == [ll.PONY, ll.SCHOLAR, ll.COP, ll.PRINCESS]
    # life.rs:1339-1348
    _, commits = eng.lookahead(ll.reconstruct("8/8/4k3/8/8/8/8/4K2R w K -"))
    assert "8/8/4k3/8/8/8/8/5RK1 b - -" in [ll.preserve(c["tree"]) for c in commits]
    # life.rs:1596-1602: fool's mate -> nothing to do
    fool = ll.reconstruct("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq -")
    assert eng.lookahead(fool)[0][1] == 0
    # life.rs:1659-1682: passing by
    _, commits = eng.lookahead(ll.reconstruct("rnbqkbnr/ppp2ppp/4p3/3pP3/8/8/PPPP1PPP/RNBQKBNR w KQkq d6 0 3"))
    ep = [c for c in commits if c["whence"] == ll.algebraic("e5")]
    assert len(ep) == 1 and ll.preserve(ep[0]["tree"]) == "rnbqkbnr/ppp2ppp/3Pp3/8/8/8/PPPP1PPP/RNBQKBNR b KQkq -"
    assert ep[0]["hospitalization"] == ll.agent(ll.BLUE, ll.SERVANT)


def test_lookahead_empty_batch_and_capacity_error(eng):
    import leafline_b200 as ll
    from leafline_b200 import _abi
    import ctypes as C

    offsets, commits = eng.lookahead(np.zeros(0, dtype=O.WORLD_DTYPE))
    assert list(offsets) == [0] and len(commits) == 0
    w = np.ascontiguousarray(ll.new_world()).reshape(1)
    offs = np.zeros(2, dtype=np.uint32)
    small = np.zeros(5, dtype=ll.COMMIT_DTYPE)
    rc = eng._lib.ll_lookahead_batch(eng._ctx, w.ctypes.data_as(C.c_void_p), 1, 0, offs.ctypes.data_as(C.c_void_p),
                                     small.ctypes.data_as(C.c_void_p), 5)
    assert rc == _abi.ERR_CAPACITY and offs[1] == 20
    assert not small.tobytes().strip(b"\0"), "nothing may be written past / into a too-small buffer"


def test_out_of_memory_is_recoverable():
    """LL_ERR_OOM leaves the context usable: a level whose growth failed is emptied, not left with a stale capacity
    (perft(9) of the opening needs a 3.2 G-node level, ~370 GB)."""
    import leafline_b200 as ll
    from leafline_b200 import _abi

    root = ll.new_world()
    with ll.Engine(0) as e:
        assert e.perft(root, 6)[0] == 119060324
        assert e.perft(root, 7)[0] == 3195901860  # sizes level 5 (4.9 M nodes)
        with pytest.raises(ll.LeaflineError) as err:
            e.perft(root, 9)
        assert err.value.code == _abi.ERR_OOM
        assert e.perft(root, 6)[0] == 119060324  # smaller trees again, on the same context
        assert e.perft(root, 7)[0] == 3195901860
        w = O.playout_batch(0x1EAF11E, 0, 64)
        assert e.horizon(w, 2).tobytes() == np.array([O.negamax(x, 2) for x in w], dtype=np.float32).tobytes()
        # a malformed ABI byte is rejected alike by the single-position and the batch paths
        bad = np.array([root, root], dtype=ll.WORLD_DTYPE)
        bad["service_eligibility"][1] = 0x1F
        for batch in (bad[1:], bad):
            with pytest.raises(ll.LeaflineError) as err:
                e.lookahead(batch)
            assert err.value.code == _abi.ERR_DOMAIN


def test_domain_errors_are_reported_not_guessed(eng):
    import leafline_b200 as ll

    bad = ll.new_world().copy()
    bad["plane"][1] |= bad["plane"][0]  # ponies overlap servants
    with pytest.raises(ll.LeaflineError) as e:
        eng.lookahead(np.array([ll.new_world(), bad], dtype=ll.WORLD_DTYPE))
    assert e.value.code == -4 and "position 1" in str(e.value)
    ghost = ll.new_world().copy()
    ghost["passing_by"] = ll.algebraic("e6")  # no Blue servant on e5: not a passing-by square
    with pytest.raises(ll.LeaflineError) as e:
        eng.perft(ghost, 2)
    assert e.value.code == -4
    assert eng.perft(ll.new_world(), 2)[0] == 400  # the context carries on


def test_literal_path_answers_positions_outside_domain_w(eng):
    """The reference answers any runes (`--from`, main.rs:199; lookahead loops over every figurehead, life.rs:678-690): two
    figureheads a team, seventeen figurines, eligibility with nothing in place.  Such positions take the literal path (the
    canonical generator with the reference's own make-the-move-and-test legality): lookahead, reckless_lookahead, perft,
    horizon values and kickoff scores must equal the oracle's, byte for byte."""
    import leafline_b200 as ll
    from positions import outside_w_worlds
    from test_hostcheck import walk

    roots = outside_w_worlds()
    worlds = np.array(walk(roots, 6, seed=4), dtype=ll.WORLD_DTYPE)
    assert len(worlds) > 60
    for nihil, ref in ((False, O.lookahead), (True, O.reckless_lookahead)):
        offsets, commits = eng.lookahead(worlds, nihilistically=nihil)  # one batch: some members inside W, some outside
        for i, w in enumerate(worlds):
            assert commits[offsets[i]:offsets[i + 1]].tobytes() == ref(w).tobytes(), (O.preserve(w), nihil)
        one = eng.lookahead(worlds[:1], nihilistically=nihil)[1]  # the single-position hand-over
        assert one.tobytes() == ref(worlds[0]).tobytes()
    for w in roots:
        for depth in (1, 2, 3):
            leaves, per_root = eng.perft(w, depth)
            want = O.perft_divide(w, depth, threads=4)
            assert leaves == int(want.sum()) and list(per_root) == list(want), (O.preserve(w), depth)
        parts = [eng.perft(w, 3, shard=(r, 3))[0] for r in range(3)]
        assert sum(parts) == eng.perft(w, 3)[0]
    for plies in (0, 1, 2):
        got = eng.horizon(worlds[:40], plies)
        want = np.array([O.negamax(w, plies) for w in worlds[:40]], dtype=np.float32)
        assert got.tobytes() == want.tobytes(), plies
    got = eng.horizon(worlds[:20], 1, quiet=2)
    assert got.tobytes() == np.array([O.negamax(w, 1, quiet=2) for w in worlds[:20]], dtype=np.float32).tobytes()
    for w in roots[:5]:
        r_got, s_got, _ = eng.kickoff(w, 3)
        r_want, s_want = O.kickoff(w, 3, False, threads=4)
        assert r_got.tobytes() == r_want.tobytes() and s_got.tobytes() == s_want.tobytes(), O.preserve(w)
        fc = eng.forecast(w, 2)
        assert sorted(float(x) for x in fc["score"]) == sorted(float(x) for x in O.kickoff(w, 2, False, threads=4)[1])
    with ll.MultiEngine([0, 0]) as g:  # a group answers such a tree whole on every rank
        g.set_shard_min_depth(perft_depth=2, kickoff_depth=3)
        assert g.perft(roots[0], 3)[0] == int(O.perft_divide(roots[0], 3, threads=4).sum())
        assert g.kickoff(roots[0], 3)[1].tobytes() == O.kickoff(roots[0], 3, False, threads=4)[1].tobytes()
    assert eng.perft(ll.new_world(), 5)[0] == 4865609  # and the fast path is back for the next call


def test_in_critical_endangerment(eng, playouts, reckless_worlds):
    import leafline_b200 as ll

    v_day = ll.reconstruct("rnb1kbnr/pppp1ppp/8/4p3/6Pq/5P2/PPPPP2P/RNBQKBNR w KQkq -")  # life.rs:1586-1594
    assert list(eng.in_critical_endangerment([v_day, v_day], [ll.ORANGE, ll.BLUE])) == [True, False]
    for worlds in (playouts[:300], reckless_worlds[:300]):
        for team in (0, 1):
            got = eng.in_critical_endangerment(worlds, team)
            want = [O.in_critical_endangerment(w, team) for w in worlds]
            assert list(got) == want


# ---------------------------------------------------------------- perft (recursive lookahead() leaf count)
# deeper entries: castling-free positions equal the published orthodox tables; the others are this engine's counts under
# the reference's rules, whose orthodox-rules probe build reproduces the published tables (profiles/r01_orthodox_probe.txt)
PERFT = [
    (OPENING, [20, 400, 8902, 197281, 4865609]),
    (POSITION_5, [44, 1486, 62416, 2104591, 90060265]),  # phantom cop (quirk 2); orthodox chess: 62 379 / 2 103 487 / 89 941 194
    (POSITION_3, [14, 191, 2812, 43238, 674624, 11030083, 178633661]),
    (POSITION_6, [46, 2079, 89890, 3894594, 164075551]),
    (KIWIPETE, [48, 2038, 97814, 4080840, 193457491]),
    (POSITION_4, [6, 258, 9221, 405001, 15117015]),
    (KIWIPETE, [48, 2038, 97814]),
    (POSITION_5, [44, 1486, 62416]),
    (POSITION_4, [6, 258, 9221]),
    (POSITION_3, [14, 191, 2812, 43238]),
    (POSITION_6, [46, 2079, 89890]),
]


@pytest.mark.parametrize("fen,counts", PERFT)
def test_perft_counts(eng, fen, counts):
    import leafline_b200 as ll

    w = ll.reconstruct(fen)
    assert eng.perft(w, 0)[0] == 1
    for depth, expected in enumerate(counts, start=1):
        leaves, per_root = eng.perft(w, depth)
        assert leaves == expected
        assert len(per_root) == counts[0] and int(per_root.sum()) == expected


def test_perft_depth_7_and_8_of_the_opening(eng):
    # perft(7) equals orthodox chess (no rule of the reference can differ before ply 8); perft(8) exceeds the orthodox
    # 84 998 978 956 by the 899 phantom-cop secret services of ply 8 -- the probe build with that ONE rule switched to
    # orthodox chess reproduces the published numbers exactly (profiles/r01_orthodox_probe.txt)
    import leafline_b200 as ll

    assert eng.perft(ll.new_world(), 7)[0] == 3195901860
    leaves, per_root = eng.perft(ll.new_world(), 8)
    assert leaves == 84998979855 and int(per_root.sum()) == leaves and len(per_root) == 20


@pytest.mark.parametrize("fen,depth", [(OPENING, 4), (KIWIPETE, 4), (POSITION_5, 4), (POSITION_4, 4)])
def test_perft_divide_matches_oracle(eng, fen, depth):
    import leafline_b200 as ll

    _, per_root = eng.perft(ll.reconstruct(fen), depth)
    assert list(per_root) == list(O.perft_divide(O.reconstruct(fen), depth, threads=16))


def test_perft_depth_6_opening_bit_exact(eng):
    import leafline_b200 as ll

    leaves, per_root = eng.perft(ll.new_world(), 6)
    assert leaves == 119060324  # BASELINE.json configs[1]
    assert int(per_root.sum()) == 119060324


def test_perft_deeper_trees_are_consistent_across_split_depths(eng):
    # size-independent property: perft(d) == sum over children of perft(d-1), at sizes the oracle cannot reach
    import leafline_b200 as ll

    for fen, depth in ((KIWIPETE, 5), (POSITION_4, 5), (POSITION_5, 5)):
        w = ll.reconstruct(fen)
        leaves, per_root = eng.perft(w, depth)
        _, commits = eng.lookahead(w)
        again = [eng.perft(c["tree"], depth - 1)[0] for c in commits]
        assert list(per_root) == again and leaves == sum(again)


def test_perft_launch_chain_equals_level_by_level_path(eng):
    # ll_perft first sizes the levels on the host (synced path), later calls run as one launch chain with
    # device-resident level sizes and order-free child placement, replayed as a CUDA graph; all must give the same divide table
    import leafline_b200 as ll

    for fen, depth in ((OPENING, 6), (KIWIPETE, 5), (POSITION_4, 5), (POSITION_3, 6), (OPENING, 3)):
        w = ll.reconstruct(fen)
        eng.force_synced(True)
        want = eng.perft(w, depth)
        for mode in (2, 0, 0):  # the chain launch by launch, then as a captured CUDA graph (capture, then replay)
            eng.force_synced(mode)
            got = eng.perft(w, depth)
            assert got[0] == want[0] and list(got[1]) == list(want[1]), (fen, depth, mode)
        for shard in ((1, 3), (0, 2)):
            eng.force_synced(True)
            want_s = eng.perft(w, depth, shard=shard)
            for mode in (2, 0, 0):
                eng.force_synced(mode)
                got_s = eng.perft(w, depth, shard=shard)
                assert got_s[0] == want_s[0] and list(got_s[1]) == list(want_s[1]), (fen, depth, shard, mode)
    # a tree that outgrows the buffers the previous calls left behind falls back and still counts right
    small = ll.reconstruct(POSITION_3)
    eng.perft(small, 3)
    assert eng.perft(ll.reconstruct(KIWIPETE), 3)[0] == 97814


def test_narrow_levels_warp_per_parent_expansion_on_many_positions(eng, playouts, reckless_worlds):
    # the launch chain expands its narrow levels with one warp per parent (k_expand_coop): every lane a piece / a fill
    # direction / a line through the figurehead.  perft(3) and perft(4) through the chain against the level-by-level
    # path (canonical generator) on named, playout and reckless-walk roots: divide tables must be identical.
    worlds = list(named_worlds()) + list(playouts[:300]) + list(reckless_worlds[:200])
    for depth, sample in ((3, worlds), (4, worlds[::7])):
        for w in sample:
            eng.force_synced(True)
            want = eng.perft(w, depth)
            eng.force_synced(False)
            got = eng.perft(w, depth)
            assert got[0] == want[0] and list(got[1]) == list(want[1]), (O.preserve(w), depth)
    eng.force_synced(False)


def test_perft_device_resident_counts(eng):
    import torch

    import leafline_b200 as ll

    out = torch.zeros(ll.PERFT_DEVICE_WORDS, dtype=torch.int64, device="cuda")
    for fen, depth, shard in ((OPENING, 5, (0, 1)), (KIWIPETE, 4, (0, 1)), (KIWIPETE, 4, (1, 2))):
        w = ll.reconstruct(fen)
        leaves, per_root = eng.perft(w, depth, shard=shard)  # also sizes the level buffers
        eng.perft_device(w, depth, out.data_ptr(), shard=shard)
        eng.synchronize()
        got = out.cpu().numpy().view(np.uint64)
        assert got[2] == 0 and got[0] == leaves and got[1] == len(per_root)
        assert list(got[8:8 + len(per_root)]) == list(per_root)
    with pytest.raises(ll.LeaflineError):
        ll.Engine(0).perft_device(ll.new_world(), 5, out.data_ptr())  # a fresh context has no level buffers yet


def test_kiwipete_prefix_subtrees_equal_the_oracle_recounts(eng):
    """configs[4] (perft(7) of Kiwipete, 3.7e11 leaves) is beyond the faithful oracle; SURVEY.md 8(d) prescribes the hierarchy:
    a seeded sample of 2-ply prefixes recounted by the oracle at sub-depth 5 (tests/golden/kiwipete_prefix_perft5.json, ~1.8e8
    leaves each, written by tests/golden/make_kiwipete_prefixes.py), the whole tree held by GPU self-consistency (bench.py config5:
    one GPU against many, the sum of the 48 root children)."""
    import json
    import os

    import leafline_b200 as ll

    with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "kiwipete_prefix_perft5.json")) as f:
        fx = json.load(f)
    assert fx["sub_depth"] == 5 and len(fx["samples"]) >= 16 and fx["n_prefixes"] == 2038
    root = ll.reconstruct(fx["position"])
    _, first = eng.lookahead(root)
    for s in fx["samples"]:
        w = ll.reconstruct(s["runes"])
        _, replies = eng.lookahead(first[s["root_move"]]["tree"])
        assert replies[s["reply"]]["tree"].tobytes() == w.tobytes()  # the sample really is that 2-ply prefix
        assert eng.perft(w, 5)[0] == s["perft5"], s


@pytest.mark.parametrize("world_size", [2, 3, 8])
def test_perft_shards_sum_to_the_whole(eng, world_size):
    import leafline_b200 as ll

    for fen, depth in ((OPENING, 5), (KIWIPETE, 4), (OPENING, 1), (OPENING, 2)):
        w = ll.reconstruct(fen)
        whole, whole_roots = eng.perft(w, depth)
        parts = [eng.perft(w, depth, shard=(r, world_size)) for r in range(world_size)]
        assert sum(p[0] for p in parts) == whole
        n = max(len(p[1]) for p in parts)
        summed = np.zeros(n, dtype=np.uint64)
        for p in parts:
            summed[: len(p[1])] += p[1]
        assert list(summed) == list(whole_roots)


# ---------------------------------------------------------------- horizon / kickoff (mind.rs:279-287, 375-448)
@pytest.mark.parametrize("plies", [0, 1, 2])
def test_horizon_matches_exact_negamax(eng, playouts, reckless_worlds, plies):
    n = {0: 200, 1: 100, 2: 24}[plies]
    for worlds in (playouts[:n], reckless_worlds[:n], named_worlds()[:n]):
        got = eng.horizon(worlds, plies)
        want = np.array([O.negamax(w, plies) for w in worlds], dtype=np.float32)
        assert np.array_equal(got, want), np.nonzero(got != want)


def test_horizon_launch_chain_equals_level_by_level(eng, playouts, reckless_worlds):
    # small batches run as ONE launch chain (device-resident level sizes, order-free placement, back-up by recorded
    # ranges); a level that outgrows its scratch capacity (4096 x ~35^2 nodes here) sends the call down the
    # level-by-level path: the values are the same every way
    big = np.concatenate([playouts] * 4)
    assert len(big) == 4096
    for worlds, plies in ((playouts[:300], 2), (reckless_worlds[:64], 3), (named_worlds(), 3), (big, 2), (big, 3),
                          (named_worlds()[:6], 4), (playouts[:1], 4), (playouts[7:8], 2)):
        eng.force_synced(False)
        chained = eng.horizon(worlds, plies)
        again = eng.horizon(worlds, plies)
        eng.force_synced(True)
        try:
            synced = eng.horizon(worlds, plies)
        finally:
            eng.force_synced(False)
        assert chained.tobytes() == synced.tobytes() == again.tobytes(), (len(worlds), plies)
    import leafline_b200 as ll

    bad = np.array([playouts[0], playouts[1]], dtype=ll.WORLD_DTYPE)
    bad[1]["plane"][0] |= np.uint64(1)  # overlapping planes
    bad[1]["plane"][7] |= np.uint64(1)
    with pytest.raises(ll.LeaflineError, match="position 1 cannot be answered"):
        eng.horizon(bad, 2)


def test_horizon_parents_with_more_records_than_a_tile_keeps(eng, playouts):
    """k_horizon keeps, per parent, the (piece / servant-step, target mask) records that have stunning, ascending or serving
    commits -- 14 of the servants' and the figurehead's, 10 of the other pieces' -- and values the quiet commits as the
    records are made.  A parent with more goes, whole, to k_horizon_overflow; values must not notice."""
    import leafline_b200 as ll

    overflowing = [
        "4k3/pppppppp/N1N1N1N1/QQQQQQQQ/8/8/8/4K3 w - -",      # twelve pieces, every one of them with somebody to stun
        "4k3/8/8/8/qqqqqqqq/n1n1n1n1/PPPPPPPP/4K3 b - -",      # the same for Blue
        "n1n1k3/1P6/8/3p1p2/4P3/8/7n/6K1 w - -",               # a servant ascending three ways (12 records), another stunning two ways, the figurehead
        "6k1/7N/8/4p3/3P1P2/8/1p6/N1N1K3 b - -",               # the same for Blue
    ]
    crowded = [  # many records, few of them loud: these stay in the tile
        "1n1n3k/2P5/BBBB4/RRRR4/NNNN4/5p1p/6P1/4K3 w - -",
        "4k3/6p1/5P1P/nnnn4/rrrr4/bbbb4/2p5/1N1N3K b - -",
        "1n1n3k/2P5/QQQQ4/QQQQ4/QQQQ4/5p1p/6P1/4K3 w - -",
        "1n1n1n2/P1P1P1P1/8/8/NNNN4/BBBB4/RRR5/4K2k w - -",
    ]
    worlds = np.array([ll.reconstruct(f) for f in overflowing + crowded], dtype=ll.WORLD_DTYPE)
    assert max(len(O.reckless_lookahead(w)) for w in worlds) >= 90
    mixed = np.concatenate([playouts[:100], worlds, playouts[100:164], worlds[::-1]])  # overflowing parents among ordinary ones
    eng.force_synced(True)  # (level by level: the overflow counter read below is that path's)
    try:
        got = eng.horizon(mixed, 1)
        assert eng.debug_horizon_overflowed() == 2 * len(overflowing)
        assert got.tobytes() == np.array([O.negamax(w, 1) for w in mixed], dtype=np.float32).tobytes()
        got = eng.horizon(playouts[:164], 1)
        assert eng.debug_horizon_overflowed() == 0
    finally:
        eng.force_synced(False)
    for plies in (1, 2):
        got = eng.horizon(mixed, plies)
        want = np.array([O.negamax(w, plies) for w in mixed], dtype=np.float32)
        assert got.tobytes() == want.tobytes(), plies
    got = eng.horizon(mixed, 1, quiet=2)
    assert got.tobytes() == np.array([O.negamax(w, 1, quiet=2) for w in mixed], dtype=np.float32).tobytes()
    for w in worlds[:4]:
        _, s_got, _ = eng.kickoff(w, 2)
        assert s_got.tobytes() == O.kickoff(w, 2, False, threads=4)[1].tobytes()


def test_horizon_tiles_with_more_children_than_the_map_holds(eng, playouts):
    """k_horizon looks every child's (parent, record, number within the record) up in a per-tile map of 3072 entries; a tile
    of 64 parents with more children than that finds them by search instead.  Whole tiles of rich positions (each within the
    24 records), ordinary ones around them: values must not notice."""
    import leafline_b200 as ll

    rich = [
        "1n1n1n2/P1P1P1P1/8/8/NNNN4/BBBB4/RRR5/4K2k w - -",   # 22 records
        "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -",
        "3k4/8/1q1q1q2/8/1q1q1q2/8/8/6K1 b - -",
        "R6R/3Q4/1Q4Q1/4Q3/2Q4Q/Q4Q2/pp1Q4/kBNN1KB1 w - -",   # 218 commits
    ]
    worlds = np.array([ll.reconstruct(f) for f in rich], dtype=ll.WORLD_DTYPE)
    counts = [len(O.reckless_lookahead(w)) for w in worlds]
    assert min(counts) >= 48 and max(counts) >= 200, counts  # 64 of any of them overflow the map
    batch = np.concatenate([playouts[:70], np.repeat(worlds, 96), playouts[70:100], np.tile(worlds, 40)])
    for plies, quiet in ((1, None), (2, None), (1, 2)):
        got = eng.horizon(batch, plies, quiet=quiet)
        want = {}
        for i, w in enumerate(batch):
            key = w.tobytes()
            if key not in want:
                want[key] = np.float32(O.negamax(w, plies, quiet=quiet) if quiet else O.negamax(w, plies))
            assert got[i].tobytes() == want[key].tobytes(), (i, plies, quiet)


def test_horizon_depth_3_on_a_few(eng, playouts):
    worlds = playouts[:4]
    got = eng.horizon(worlds, 3)
    want = np.array([O.negamax(w, 3) for w in worlds], dtype=np.float32)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("plies,quiet", [(0, 1), (0, 2), (1, 1), (1, 3), (2, 2), (0, 0), (1, 0)])
def test_horizon_with_quiescence_extension(eng, playouts, reckless_worlds, plies, quiet):
    # mind.rs:284-301: below depth 0 only stunning commits are searched and every node may stand pat
    n = 40 if plies + quiet <= 3 else 10
    for worlds in (playouts[100:100 + n], reckless_worlds[50:50 + n], named_worlds()[:n]):
        got = eng.horizon(worlds, plies, quiet=quiet)
        want = np.array([O.negamax(w, plies, quiet) for w in worlds], dtype=np.float32)
        assert np.array_equal(got, want), (plies, quiet, np.nonzero(got != want))


def test_kickoff_with_quiescence_extension(eng):
    import leafline_b200 as ll

    for fen, depth, quiet in ((KIWIPETE, 2, 2), (POSITION_5, 2, 1), (OPENING, 3, 2), ("8/q1P1k/8/8/8/8/6PP/7K w - -", 3, 3)):
        w = ll.reconstruct(fen)
        roots, scores, scored = eng.kickoff(w, depth, quiet=quiet)
        want_roots, want_scores = O.kickoff(w, depth, False, threads=8, quiet=quiet)
        assert roots.tobytes() == want_roots.tobytes()
        assert np.array_equal(scores, want_scores), (fen, depth, quiet)


def check_kickoff(eng, world, depth, nihilistically):
    roots, scores, scored = eng.kickoff(world, depth, nihilistically=nihilistically)
    want_roots, want_scores = O.kickoff(world, depth, nihilistically, threads=8)
    assert roots.tobytes() == want_roots.tobytes()
    assert np.array_equal(scores, want_scores)
    assert scored > 0
    return roots, scores


def test_kickoff_known_answers_of_the_reference(eng):
    import leafline_b200 as ll

    # mind.rs:662-675: under-promotion to a pony
    ws = ll.reconstruct("8/q1P1k/8/8/8/8/6PP/7K w - -")
    roots, scores = check_kickoff(eng, ws, 3, True)
    best = int(np.argmax(scores))
    assert scores[best] > 0 and ll.preserve(roots[best]["tree"]) == "2N5/q3k3/8/8/8/8/6PP/7K b - -"
    assert float(scores[best]) == 4.69921875
    # mind.rs:677-734: take the pony, both colours
    for fen in ("1p6/8/2N5/8/6p1/2B5/8/n7 w - -", "1P6/8/2n5/8/6P1/2b5/8/N7 b - -"):
        roots, scores = check_kickoff(eng, ll.reconstruct(fen), 2, True)
        assert roots[int(np.argmax(scores))]["whither"] == ll.locale(0, 0)


def oracle_variation_values(world, depth, quiet):
    """value of every node along a variation must equal the exact negamax value of that node (sign-alternating)"""
    return O.negamax(world, depth, quiet)


@pytest.mark.parametrize("fen,depth,quiet", [(KIWIPETE, 3, None), ("8/q1P1k/8/8/8/8/6PP/7K w - -", 3, None), (POSITION_5, 2, 2),
                                             (OPENING, 4, None), (POSITION_4, 3, 1)])
def test_forecasts_are_sorted_and_their_variations_are_principal(eng, fen, depth, quiet):
    import leafline_b200 as ll

    w = ll.reconstruct(fen)
    fc = eng.forecast(w, depth, quiet=quiet)
    roots, scores, _ = eng.kickoff(w, depth, quiet=quiet)
    assert len(fc) == len(roots)
    assert list(fc["score"]) == sorted(scores, reverse=True)  # mind.rs:436
    # stable: equal scores keep the canonical order
    order = sorted(range(len(scores)), key=lambda i: -scores[i])
    assert [bytes(f["commit"].tobytes()) for f in fc] == [roots[i].tobytes() for i in order]
    for f in fc[:6]:
        n = int(f["variation_len"])
        assert 1 <= n <= depth + (quiet or 0)
        assert f["variation"][0]["whence"] == f["commit"]["whence"] and f["variation"][0]["whither"] == f["commit"]["whither"]
        # replay the line with the oracle: every step must be a reckless commit of the node, and the value must carry along it
        node, remaining, value = f["commit"]["tree"].copy(), depth - 1, -np.float32(f["score"])
        for step in f["variation"][1:n]:
            assert O.negamax(node, remaining, quiet) == value
            kids = [c for c in O.reckless_lookahead(node) if c["whence"] == step["whence"] and c["whither"] == step["whither"] and c["job"] == step["job"]]
            assert kids, "variation step is not a move of the node"
            vals = [-O.negamax(c["tree"], remaining - 1, quiet) for c in kids]
            k = int(np.argmax([v == value for v in vals]))
            assert vals[k] == value, "variation step does not carry the node's value"
            node, remaining, value = kids[k]["tree"].copy(), remaining - 1, -value
        assert O.negamax(node, remaining, quiet) == value


@pytest.mark.parametrize("fen,depth,device_plies,quiet", [(KIWIPETE, 4, 2, None), (OPENING, 5, 3, None), (POSITION_5, 4, 1, None),
                                                          (POSITION_4, 4, 2, 1), (KIWIPETE, 3, 0, None), ("8/q1P1k/8/8/8/8/6PP/7K w - -", 5, 2, None)])
def test_host_alpha_beta_driver_equals_full_width_search(eng, fen, depth, device_plies, quiet):
    # mind.rs:271-364 on the host over device frontier batches: cut-offs prune work, never the root scores
    import leafline_b200 as ll

    w = ll.reconstruct(fen)
    pruned, expanded, handed_off = eng.alpha_beta_forecast(w, depth, device_plies=device_plies, quiet=quiet)
    full = eng.forecast(w, depth, quiet=quiet)
    assert pruned["score"].tobytes() == full["score"].tobytes()
    assert pruned["commit"].tobytes() == full["commit"].tobytes()
    assert handed_off > 0
    if depth - device_plies >= 3:  # cut-offs happened: fewer frontier nodes than the full-width tree has at that ply
        frontier = np.array([w], dtype=ll.WORLD_DTYPE)
        frontier = np.ascontiguousarray(eng.lookahead(frontier)[1]["tree"])
        for _ in range(depth - device_plies - 1):
            frontier = np.ascontiguousarray(eng.reckless_lookahead(frontier)[1]["tree"])
        assert handed_off < len(frontier)


def test_host_alpha_beta_memory_bank_is_sound(eng):
    # mind.rs:312-344 made sound (whole-position keys, exact/lower/upper kinds): transpositions are answered from the
    # bank and the root scores are still those of the full-width search, bit for bit; a root nearer than the hand-off
    # depth is searched to ITS depth
    import leafline_b200 as ll

    for fen, depth, device_plies in ((OPENING, 6, 2), (OPENING, 6, 3), (KIWIPETE, 5, 2), ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -", 7, 2),
                                     (OPENING, 3, 4), (KIWIPETE, 2, 4)):
        w = ll.reconstruct(fen)
        pruned, expanded, handed_off = eng.alpha_beta_forecast(w, depth, device_plies=device_plies)
        full = eng.forecast(w, depth)
        assert pruned["score"].tobytes() == full["score"].tobytes(), (fen, depth, device_plies)
        assert pruned["commit"].tobytes() == full["commit"].tobytes()
        if depth - device_plies >= 4:
            assert eng.last_search_stats["values_remembered"] > 0


def test_host_alpha_beta_variations_memory_budget_and_group(eng):
    """mind.rs:208-224, 367-372, 399-414: the driver returns real principal variations (every line, played out, reaches a
    position whose static value is the forecast's score), forgets under a byte budget without changing a score, and
    searches the root moves concurrently over a group of GPUs with the same results."""
    import leafline_b200 as ll

    for fen, depth, device_plies in ((OPENING, 6, 3), (KIWIPETE, 5, 2), ("8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -", 7, 3)):
        w = ll.reconstruct(fen)
        full = eng.forecast(w, depth)
        pruned, _, _ = eng.alpha_beta_forecast(w, depth, device_plies=device_plies)
        assert pruned["score"].tobytes() == full["score"].tobytes()
        for f in pruned[:6]:
            n = int(f["variation_len"])
            assert n == depth, (fen, n)  # a full line: nobody ran out of moves on these
            cur = w
            for p in f["variation"][:n]:  # play the line with the oracle's generator (ply 0 legal, below it reckless)
                kids = O.lookahead(cur) if cur is w else O.reckless_lookahead(cur)
                hit = [c for c in kids if (c["team"], c["job"], c["whence"], c["whither"]) == (p["team"], p["job"], p["whence"], p["whither"])]
                assert hit, (fen, ll.pagan_variation(f))
                # (an ascension's four commits share a patch: any of them may be the one; take the one that keeps the score)
                cur = min(hit, key=lambda c: abs(float(O.score(c["tree"])) * (1 if w["initiative"] == 0 else -1) - float(f["score"])))["tree"] if len(hit) > 1 else hit[0]["tree"]
            leaf = float(O.score(cur)) * (1.0 if int(w["initiative"]) == 0 else -1.0)
            assert np.float32(leaf) == f["score"], (fen, ll.pagan_variation(f), leaf, float(f["score"]))
    # one searcher thread (a single depth-first walk) and thirty-two (one per root move, requests combined): the same scores
    w = ll.reconstruct(KIWIPETE)
    eng.set_search_threads(1)
    alone, _, _ = eng.alpha_beta_forecast(w, 5, device_plies=2)
    eng.set_search_threads(32)
    together, _, _ = eng.alpha_beta_forecast(w, 5, device_plies=2)
    assert alone["score"].tobytes() == together["score"].tobytes() and alone["commit"].tobytes() == together["commit"].tobytes()
    # a tiny memory bank forgets (least recently used first) and the scores do not move
    w = ll.new_world()
    full = eng.forecast(w, 6)
    eng.set_deja_vu_bound(1e-4)  # ~300 entries
    tight, _, _ = eng.alpha_beta_forecast(w, 6, device_plies=2)
    eng.set_deja_vu_bound(0.25)
    assert tight["score"].tobytes() == full["score"].tobytes()
    # the group: root moves dealt to the members, one max-reduction gathers scores and lines
    with ll.MultiEngine([0, 0, 0]) as g:
        for fen, depth, device_plies in ((OPENING, 6, 3), (KIWIPETE, 5, 2)):
            w = ll.reconstruct(fen)
            one, _, handed_one = eng.alpha_beta_forecast(w, depth, device_plies=device_plies)
            many, _, handed_many = g.alpha_beta_forecast(w, depth, device_plies=device_plies)
            assert many["score"].tobytes() == one["score"].tobytes() and many["commit"].tobytes() == one["commit"].tobytes()
            assert [int(x) for x in many["variation_len"]] == [int(x) for x in one["variation_len"]]
            assert handed_many > 0


def test_kickoff_family(eng):
    import leafline_b200 as ll

    w = ll.reconstruct("8/q1P1k/8/8/8/8/6PP/7K w - -")  # mind.rs:662-675
    fc = eng.fixed_depth_sequence_kickoff(w, [1, 2, 3], nihilistically=True)
    assert ll.preserve(fc[0]["commit"]["tree"]) == "2N5/q3k3/8/8/8/8/6PP/7K b - -" and fc[0]["score"] > 0
    assert ll.pagan_variation(fc[0]).startswith("c7c8")
    best, depth = eng.iterative_deepening_kickoff(ll.new_world(), 0.5)
    assert depth >= 3 and len(best) == 20
    assert np.array_equal(np.sort(best["score"])[::-1], best["score"])
    if depth <= 7:  # a full-width iteration: forecasts with variations
        assert best.tobytes() == eng.forecast(ll.new_world(), depth).tobytes()
    else:  # on a warm context depth 8 (alpha-beta driver) can fit half a second
        same, _, _ = eng.alpha_beta_forecast(ll.new_world(), depth)
        assert best["score"].tobytes() == same["score"].tobytes() and best["commit"].tobytes() == same["commit"].tobytes()
    # past the depth whose full-width tree fits HBM the host alpha-beta driver carries on, under the same deadline
    import time

    t0 = time.perf_counter()
    deeper, depth = eng.iterative_deepening_kickoff(ll.new_world(), 4.0)
    assert time.perf_counter() - t0 < 5.5
    assert depth >= 8 and len(deeper) == 20
    same, _, _ = eng.alpha_beta_forecast(ll.new_world(), depth, device_plies=3 if depth <= 8 else 5)
    assert deeper["score"].tobytes() == same["score"].tobytes() and deeper["commit"].tobytes() == same["commit"].tobytes()
    # a wider tree: the budget is kept and the deepest finished iteration is what comes back
    w = ll.reconstruct(KIWIPETE)
    t0 = time.perf_counter()
    kept, depth = eng.iterative_deepening_kickoff(w, 1.2)
    assert time.perf_counter() - t0 < 2.5 and depth >= 4
    if depth <= 6:
        assert kept.tobytes() == eng.forecast(w, depth).tobytes()
    with pytest.raises(ll.LeaflineError):
        eng.fixed_depth_sequence_kickoff(w, [])


@pytest.mark.parametrize("depth", [1, 2, 3])
def test_kickoff_opening_matches_oracle(eng, depth):
    import leafline_b200 as ll

    _, scores = check_kickoff(eng, ll.new_world(), depth, False)
    top = {1: -0.4, 2: 0.5, 3: -0.4}[depth]
    assert abs(float(scores.max()) - top) < 1e-6


def test_kickoff_depth_4_kiwipete_matches_oracle(eng):
    import leafline_b200 as ll

    check_kickoff(eng, ll.reconstruct(KIWIPETE), 3, False)
    check_kickoff(eng, ll.reconstruct(POSITION_5), 3, True)


def test_kickoff_depth_5_and_6_match_the_oracle(eng):
    # BASELINE.json configs[3]: depth-5 search from the opening (and one ply deeper), per-root scores equal to TT-free
    # exact negamax (the oracle prunes with alpha-beta, which is exact at the root; the device searches full width)
    import leafline_b200 as ll

    for fen, depth in ((OPENING, 5), (OPENING, 6), (KIWIPETE, 4)):
        w = ll.reconstruct(fen)
        roots, scores, scored = eng.kickoff(w, depth)
        want_roots, want_scores = O.kickoff(w, depth, False, threads=16)
        assert roots.tobytes() == want_roots.tobytes()
        assert np.array_equal(scores, want_scores), (fen, depth)
        pruned, _, _ = eng.alpha_beta_forecast(w, depth, device_plies=3)
        assert sorted(pruned["score"]) == sorted(want_scores)


def test_kickoff_launch_chain_equals_level_by_level(eng):
    # ll_kickoff (scores only, no quiet extension) searches the levels below the root moves as one launch chain; the first
    # call at a new size may outgrow a scratch level and fall back, the second one stays on the chain: same scores, same
    # number of leaves scored, sharded or not
    import leafline_b200 as ll

    for fen, depth in ((OPENING, 3), (OPENING, 4), (OPENING, 5), (OPENING, 6), (KIWIPETE, 4), (KIWIPETE, 5), (POSITION_5, 4)):
        w = ll.reconstruct(fen)
        for shard in (None, (1, 3)):
            kw = {} if shard is None else {"shard": shard}
            first = eng.kickoff(w, depth, **kw)
            second = eng.kickoff(w, depth, **kw)
            eng.force_synced(True)
            try:
                synced = eng.kickoff(w, depth, **kw)
            finally:
                eng.force_synced(False)
            for got in (first, second):
                assert got[0].tobytes() == synced[0].tobytes()
                assert got[1].tobytes() == synced[1].tobytes(), (fen, depth, shard)
                assert got[2] == synced[2]


@pytest.mark.parametrize("world_size", [2, 4])
def test_kickoff_shards_gather_to_the_whole(eng, world_size):
    import leafline_b200 as ll

    w = ll.reconstruct(KIWIPETE)
    _, whole, _ = eng.kickoff(w, 3)
    gathered = np.full(len(whole), -np.inf, dtype=np.float32)
    for r in range(world_size):
        _, part, _ = eng.kickoff(w, 3, shard=(r, world_size))
        gathered = np.maximum(gathered, part)
    assert np.array_equal(gathered, whole)


# ---------------------------------------------------------------- seeded playout set (config 3)
def test_device_playouts_match_oracle(eng, playouts):
    soa = eng.playout_soa(1024)
    eng.synchronize()
    got = eng.soa_download(soa)
    assert got.tobytes() == playouts.tobytes()
    tail = eng.playout_soa(64, first_index=5000)
    eng.synchronize()
    assert eng.soa_download(tail).tobytes() == O.playout_batch(0x1EAF11E, 5000, 64).tobytes()
    eng.soa_free(soa)
    eng.soa_free(tail)


def test_full_size_score_batch_properties(eng):
    # BASELINE.json configs[2] at full size (2^24 positions): properties the oracle need not be run for
    import torch

    n = 1 << 24
    soa = eng.playout_soa(n)
    out = torch.empty(n, dtype=torch.float32, device="cuda")
    eng.score_soa(soa, out.data_ptr())
    eng.synchronize()
    scores = out.cpu().numpy()
    assert np.isfinite(scores).all()
    # prefix against the oracle, bit-exact
    k = 4096
    worlds = eng.soa_download(soa, 0, k)
    assert worlds.tobytes() == O.playout_batch(0x1EAF11E, 0, k).tobytes()
    assert scores[:k].tobytes() == O.score_batch(worlds).tobytes()
    # ALL 2^24 scores against the oracle, bit for bit (tolerance of the north star: 1e-5 relative)
    chunk = 1 << 21
    for first in range(0, n, chunk):
        worlds = eng.soa_download(soa, first, chunk)
        want = O.score_batch(worlds)
        np.testing.assert_allclose(scores[first:first + chunk], want, rtol=SCORE_RTOL, atol=0)
        assert scores[first:first + chunk].tobytes() == want.tobytes(), first
    # idempotence: scoring again gives identical bits
    out2 = torch.empty_like(out)
    eng.score_soa(soa, out2.data_ptr())
    eng.synchronize()
    assert torch.equal(out, out2)
    eng.soa_free(soa)


def test_full_size_movegen_self_consistency(eng):
    # 2^22 seeded playout positions (and the depth-4 frontier of Kiwipete, rich in checks, pins, ascensions and
    # secret service): set-wise counter == generator with legality by masks == make-the-move-and-test definition
    import torch

    n = 1 << 22
    soa = eng.playout_soa(n)
    bad, first, legal, reckless = eng.selfcheck_soa(soa)
    assert bad == 0, (bad, first, O.preserve(eng.soa_download(soa, first, 1)[0]) if first is not None else None)
    assert 20 * n < legal <= reckless < 60 * n
    k = 2048  # the totals are anchored to the oracle on a prefix
    worlds = eng.soa_download(soa, 0, k)
    sub = eng.soa_upload(worlds)
    _, _, legal_k, reckless_k = eng.selfcheck_soa(sub)
    assert legal_k == sum(len(O.lookahead(w)) for w in worlds[:k]) and reckless_k == sum(len(O.reckless_lookahead(w)) for w in worlds[:k])
    eng.soa_free(sub)
    eng.soa_free(soa)
    import leafline_b200 as ll

    frontier = np.array([ll.reconstruct(KIWIPETE)], dtype=ll.WORLD_DTYPE)
    for _ in range(3):
        _, commits = eng.reckless_lookahead(frontier)
        frontier = np.ascontiguousarray(commits["tree"])
    soa = eng.soa_upload(frontier)
    bad, first, legal, reckless = eng.selfcheck_soa(soa)
    assert bad == 0 and len(frontier) > 90000, (bad, first)
    eng.soa_free(soa)


# ---------------------------------------------------------------- --correspond wire format (main.rs:182-244)
def test_correspondence_wire_format(eng):
    import json

    from leafline_b200 import correspondence as co

    # main.rs:461-469: the reference's own end-to-end test
    assert co.correspondence(eng, "R6k/6pp/8/8/8/8/8/8 b - -", co.Depth(2)) == '{"the_triumphant":"Orange"}'
    # stalemate of the side to move -> nobody triumphs
    assert co.correspondence(eng, "7k/5Q2/6K1/8/8/8/8/8 b - -", co.Depth(1)) == '{"the_triumphant":null}'
    card = co.correspondence(eng, "8/q1P1k/8/8/8/8/6PP/7K w - -", co.DepthSequence.from_depiction("1,2,3"))
    d = json.loads(card)
    assert list(d) == ["world", "patch", "hospitalization", "thinking_time", "depth", "counterreplies", "rosetta_stone"]
    assert d["world"] == "2N5/q3k3/8/8/8/8/6PP/7K b - -" and d["depth"] == 3 and d["hospitalization"] is None
    assert d["patch"] == {"star": {"team": "Orange", "job_description": "Servant"}, "whence": {"rank": 6, "file": 2}, "whither": {"rank": 7, "file": 2}}
    assert d["rosetta_stone"] == "c8" and len(d["counterreplies"]) == len(O.lookahead(O.reconstruct(d["world"])))
    assert ", " not in card and "\": " not in card  # compact, like rustc-serialize
    # a capture reports its patient; Seconds bound runs the iterative deepening
    d = json.loads(co.correspondence(eng, "1p6/8/2N5/8/6p1/2B5/8/n7 w - -", co.Depth(2)))
    assert d["hospitalization"] == {"team": "Blue", "job_description": "Pony"} and d["rosetta_stone"] == "Ba1"
    d = json.loads(co.correspondence(eng, "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -", co.Seconds(1)))
    assert d["depth"] >= 3 and len(d["counterreplies"]) == 20
    # a depth whose full-width tree does not fit HBM goes to the host alpha-beta driver: same card
    d = json.loads(co.correspondence(eng, "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -", co.Depth(8)))
    assert d["depth"] == 8 and len(d["counterreplies"]) == 20 and d["patch"]["star"]["team"] == "Orange"


def test_uci_daemon_and_command_line(eng):
    import io
    import json
    import subprocess
    import sys

    from leafline_b200 import uci

    # uci.rs:13-104: the same dialogue the reference holds
    script = "uci\nisready\nucinewgame\nposition startpos moves e2e4\ngo depth 2\nposition startpos moves e2e4 x g1f3\ngo wtime 4000 btime 4000 movestogo 2\nquit\n"
    out = io.StringIO()
    uci.dæmon(engine=eng, stdin=io.StringIO(script), stdout=out)
    lines = out.getvalue().splitlines()
    assert lines[0].startswith("id name Leafline") and lines[2] == "uciok" and lines[3] == "readyok"
    best = [l for l in lines if l.startswith("bestmove ")]
    assert len(best) == 2 and all(len(b.split()[1]) == 4 for b in best)
    assert any(l.startswith("info depth ") and " score cp " in l for l in lines)
    # the first reply answers 1.e4 at depth 2 with the move exact negamax ranks first (stable order)
    import leafline_b200 as ll

    after_e4 = [c for c in eng.lookahead(ll.new_world())[1] if ll.to_algebraic(c["whence"]) + ll.to_algebraic(c["whither"]) == "e2e4"][0]["tree"]
    top = eng.forecast(after_e4, 2)[0]["commit"]
    assert best[0] == "bestmove " + ll.to_algebraic(top["whence"]) + ll.to_algebraic(top["whither"])
    # main.rs: `leafline --correspond --depth 2 --from <runes>` as the web client runs it (main.rs:461-469 for the answer)
    root = __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__)))
    run = subprocess.run([sys.executable, "-m", "leafline_b200", "--correspond", "--depth", "2", "--from", "R6k/6pp/8/8/8/8/8/8 b - -"],
                         capture_output=True, text=True, cwd=root, timeout=300)
    assert run.returncode == 0 and run.stdout.strip() == '{"the_triumphant":"Orange"}', run.stderr
    run = subprocess.run([sys.executable, "-m", "leafline_b200", "--correspond", "--depth", "2", "--seconds", "1", "--from", "x"],
                         capture_output=True, text=True, cwd=root, timeout=300)
    assert run.returncode != 0 and "more than one of" in run.stderr


def test_one_context_per_host_thread():
    # the reference calls the leaf layer from one thread per root move (mind.rs:405); the ABI's contract is one ll_ctx per
    # host thread -- four threads, four contexts, concurrent perft / lookahead / score (ctypes releases the GIL)
    import threading

    import leafline_b200 as ll

    fens = [OPENING, KIWIPETE, POSITION_4, POSITION_5]
    with ll.Engine(0) as e0:
        want = {f: e0.perft(ll.reconstruct(f), 5 if f == OPENING else 4) for f in fens}
    errors = []

    def work(fen):
        try:
            with ll.Engine(0) as e:
                w = ll.reconstruct(fen)
                for _ in range(6):
                    got = e.perft(w, 5 if fen == OPENING else 4)
                    assert got[0] == want[fen][0] and list(got[1]) == list(want[fen][1])
                    offsets, commits = e.lookahead(w)
                    assert offsets[1] == len(want[fen][1])
                    assert e.score(commits["tree"]).tobytes() == O.score_batch(commits["tree"]).tobytes()
        except Exception as exc:  # noqa: BLE001
            errors.append((fen, repr(exc)))

    threads = [threading.Thread(target=work, args=(f,)) for f in fens]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors
