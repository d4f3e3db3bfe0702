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
