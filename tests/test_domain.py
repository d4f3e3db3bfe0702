"""Host-only: the well-formed domain W is closed under the reference's own move generator, including
the reckless lines (figurehead captures, phantom castling) the horizon search visits."""
import numpy as np

import oracle as O
from positions import NAMED, is_well_formed, named_worlds, reckless_walk_worlds


def test_named_positions_are_well_formed():
    for fen, w in zip(NAMED, named_worlds()):
        assert is_well_formed(w), fen


def test_reckless_walks_stay_well_formed():
    worlds = reckless_walk_worlds(seed=11, games=40, plies=60)
    assert len(worlds) > 1000
    assert all(is_well_formed(w) for w in worlds)
    # the walk really does reach the quirky corners
    no_king = sum(1 for w in worlds if int(w["plane"][5]) == 0 or int(w["plane"][11]) == 0)
    assert no_king > 20


def test_playout_positions_are_well_formed():
    worlds = O.playout_batch(0x1EAF11E, 0, 200)
    assert all(is_well_formed(w) for w in worlds)
    lens = [len(O.lookahead(w)) for w in worlds[:50]]
    assert 15 < np.mean(lens) < 50


def test_the_tightened_domain_is_closed_under_secret_service_with_a_phantom_cop():
    """A team of sixteen with an eligibility bit and no cop on that corner would grow a seventeenth figurine by secret service
    (life.rs:621-628): such positions belong to the literal domain, and nothing inside W can reach one."""
    from positions import OUTSIDE_W

    for fen in OUTSIDE_W:
        assert not is_well_formed(O.reconstruct(fen)), fen
    # sixteen figurines, the cop standing on its corner: inside W, and so are all its children
    w = O.reconstruct("4k3/8/8/8/8/NN6/PPPPPPPP/RNBQK2R w K -")
    assert is_well_formed(w)
    for c in O.reckless_lookahead(w):
        assert is_well_formed(c["tree"])
    # a phantom cop appears only where a seat is free: fifteen figurines, east eligibility, h1 empty
    w = O.reconstruct("4k3/8/8/8/8/NN6/PPPPPPPP/RNBQK3 w K -")
    assert is_well_formed(w)
    kids = O.reckless_lookahead(w)
    assert any(bin(int(np.bitwise_or.reduce(c["tree"]["plane"][:6]))).count("1") == 16 for c in kids)
    assert all(is_well_formed(c["tree"]) for c in kids)
