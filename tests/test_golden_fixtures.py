"""The committed golden vectors (tests/golden/golden.json, made by tests/golden/make_golden.py) against
the oracle on the CPU box and against the CUDA path (through the C ABI) on the GPU box."""
import hashlib
import json
import os

import numpy as np
import pytest

import oracle as O

GOLDEN = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden.json")))


def test_oracle_reproduces_the_golden_vectors():
    for g in GOLDEN["positions"]:
        w = O.reconstruct(g["fen"])
        assert O.preserve(w) == g["preserved"]
        assert int(np.float32(O.score(w)).view(np.uint32)) == g["score_bits"], g["fen"]
        legal, reckless = O.lookahead(w), O.reckless_lookahead(w)
        assert (len(legal), len(reckless)) == (g["legal"], g["reckless"]), g["fen"]
        assert hashlib.sha256(legal.tobytes()).hexdigest() == g["legal_sha256"], g["fen"]
        assert hashlib.sha256(reckless.tobytes()).hexdigest() == g["reckless_sha256"], g["fen"]
        assert [O.perft(w, d + 1) for d in range(len(g["perft"]))] == g["perft"], g["fen"]
        assert [int(np.float32(O.negamax(w, d)).view(np.uint32)) for d in range(3)] == g["negamax_bits"], g["fen"]
    p = GOLDEN["playouts"]
    worlds = O.playout_batch(GOLDEN["seed"], 0, p["n"])
    assert hashlib.sha256(worlds.tobytes()).hexdigest() == p["worlds_sha256"]
    assert [O.preserve(w) for w in worlds[:16]] == p["fens"]
    assert [int(x) for x in O.score_batch(worlds).view(np.uint32)] == p["score_bits"]


@pytest.mark.gpu
def test_cuda_path_reproduces_the_golden_vectors():
    import __graft_entry__ as entry

    entry.build()
    import leafline_b200 as ll

    with ll.Engine(0) as eng:
        worlds = np.array([ll.reconstruct(g["fen"]) for g in GOLDEN["positions"]], dtype=ll.WORLD_DTYPE)
        scores = eng.score(worlds)
        assert [int(x) for x in scores.view(np.uint32)] == [g["score_bits"] for g in GOLDEN["positions"]]
        for nihil, key in ((False, "legal"), (True, "reckless")):
            offsets, commits = eng.lookahead(worlds, nihilistically=nihil)
            for i, g in enumerate(GOLDEN["positions"]):
                mine = commits[offsets[i]:offsets[i + 1]]
                assert len(mine) == g[key], g["fen"]
                assert hashlib.sha256(mine.tobytes()).hexdigest() == g[key + "_sha256"], g["fen"]
        for i, g in enumerate(GOLDEN["positions"]):
            assert [eng.perft(worlds[i], d + 1)[0] for d in range(len(g["perft"]))] == g["perft"], g["fen"]
        for plies in range(3):
            vals = eng.horizon(worlds, plies)
            want = np.array([g["negamax_bits"][plies] for g in GOLDEN["positions"]], dtype=np.uint32).view(np.float32)
            assert np.array_equal(vals, want), plies
        p = GOLDEN["playouts"]
        soa = eng.playout_soa(p["n"], seed=GOLDEN["seed"])
        eng.synchronize()
        got = eng.soa_download(soa)
        eng.soa_free(soa)
        assert hashlib.sha256(got.tobytes()).hexdigest() == p["worlds_sha256"]
        assert [int(x) for x in eng.score(got).view(np.uint32)] == p["score_bits"]
        assert list(eng.lookahead_counts(got)) == p["legal_counts"]
        assert list(eng.lookahead_counts(got, nihilistically=True)) == p["reckless_counts"]
