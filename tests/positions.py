"""Shared fixtures: named positions of the reference's tests / SURVEY.md and seeded position sets."""
import random

import numpy as np

import oracle as O

OPENING = "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -"
VISION = "3q1rk1/2R1bppp/pP2p3/N2b4/1r6/4BP2/1P1Q2PP/R5K1 b - -"  # life.rs:1141
KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
POSITION_3 = "8/2p5/3p4/KP5r/1R3p1k/8/4P1P1/8 w - -"
POSITION_4 = "r3k2r/Pppp1ppp/1b3nbN/nP6/BBP1P3/q4N2/Pp1P2PP/R2Q1RK1 w kq -"
POSITION_5 = "rnbq1k1r/pp1Pbppp/2p5/8/2B5/8/PPP1NnPP/RNBQK2R w KQ -"
POSITION_6 = "r4rk1/1pp1qppp/p1np1n2/2b1p1B1/2B1P1b1/P1NP1N2/1PP1QPPP/R4RK1 w - -"

# every position string that appears in the reference's own tests for this path, plus SURVEY.md 8(a) quirks
NAMED = [
    OPENING, VISION, KIWIPETE, POSITION_3, POSITION_4, POSITION_5, POSITION_6,
    "8/8/4k3/8/8/8/8/4K2R w K -", "8/8/4k3/8/8/8/8/R3K2R w KQ -", "8/8/4k3/8/8/8/8/R3KN1R w Q -",
    "8/8/4k3/8/8/4b3/8/R3KN1R w Q -", "8/8/4k3/8/b7/8/8/R3KN1R w Q - 0 1", "8/8/4k3/8/4r3/8/8/4K2R w K -",
    "rnbqkbnr/pppppppp/8/8/8/5N2/PPPPBPPP/RNBQK2R w KQkq -",
    "rnbqkbnr/ppp2ppp/4p3/3pP3/8/8/PPPP1PPP/RNBQKBNR w KQkq d6 0 3",
    "rnbqk2r/p1pp1ppp/1p5n/8/1bPPpP2/6PP/PP1BP3/RN1QKBNR b KQkq f3 0 6",
    "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3",
    "rnbqkbnr/pp1ppppp/8/2p5/4P3/8/PPPP1PPP/RNBQKBNR w KQkq c6",
    "8/q1P1k/8/8/8/8/6PP/7K w - -", "k7/pp6/8/8/8/P7/P7/K7 w - -", "k7/pp6/8/8/8/8/PP6/K7 w - -",
    "R6k/6pp/8/8/8/8/8/8 b - -", "7K/r7/1r6/8/8/8/8/7k b - -",
    "4k3/8/8/8/8/8/8/4K2n w K -", "4k3/8/8/7R/8/8/8/4K2R w K -", "4k3/8/8/8/8/8/8/7R w K -",
    "rnbqkbnr/ppp2ppp/4p3/3p4/2PP4/8/PP2PPPP/RNBQKBNR w KQkq - 0 3",
    "r2q1rk1/2p1ppbp/1p4p1/p2p1b2/NnPPn3/PP2PN1P/1B2BPP1/R2Q1RK1 b - - 0 14",
    "r3k2r/8/8/8/8/8/8/R3K2R b KQkq -", "r3k2r/8/8/8/8/8/6n1/R3K2R w KQkq -", "4k3/8/8/8/8/8/8/R3q2R w KQ -",
    "8/P6k/8/8/8/8/p6K/8 w - -", "8/P6k/8/8/8/8/p6K/8 b - -", "1n6/P6k/8/8/8/8/p6K/1N6 b - -",
    "8/8/8/2k5/3Pp3/8/8/4K3 b - d3", "8/8/8/8/k2Pp2R/8/8/4K3 b - d3", "4k3/8/8/8/8/8/8/4K3 w - -",
]


def named_worlds():
    return np.array([O.reconstruct(f) for f in NAMED], dtype=O.WORLD_DTYPE)


def reckless_walk_worlds(seed, games, plies):
    """States reached by random PSEUDO-legal play (figureheads get captured, phantom castling happens):
    the domain the horizon search actually visits below a figurehead capture."""
    rng = random.Random(seed)
    out = []
    for _ in range(games):
        w = O.new_world() if rng.random() < 0.5 else O.reconstruct(rng.choice([KIWIPETE, POSITION_4, POSITION_5]))
        for _ in range(plies):
            moves = O.reckless_lookahead(w)
            if len(moves) == 0:
                break
            # prefer captures sometimes so that figureheads and cops do get taken
            caps = [m for m in moves if m["hospitalization"] != O.NONE]
            m = rng.choice(caps) if caps and rng.random() < 0.5 else moves[rng.randrange(len(moves))]
            w = m["tree"].copy()
            out.append(w)
    return np.array(out, dtype=O.WORLD_DTYPE)


# positions OUTSIDE domain W that the reference answers all the same (`--from` takes any runes, main.rs:199; lookahead loops over
# every figurehead, life.rs:678-690): the literal path's test set
OUTSIDE_W = [
    "4k3/8/8/8/8/8/3K4/4K2R w K -",            # two Orange figureheads, one of them on its home square: secret service live
    "4k2k/8/8/8/8/8/r7/K3K3 w - -",            # two figureheads either side; the a1 one is boxed in
    "k3k3/8/8/3Q4/8/8/8/4K3 b - -",            # the mover has two figureheads, both endangered by one princess
    "kk6/8/8/8/8/8/PPPPPPPP/RNBQKBNR w KQ -",  # two Blue figureheads
    "4k3/pppppppp/8/8/8/N7/PPPPPPPP/RNBQKBNR w KQ -",  # seventeen Orange figurines
    "rnbqkbnr/pppppppp/n6n/8/8/8/8/4K3 b kq -",        # eighteen Blue figurines
    "4k3/8/8/8/8/8/8/R2K3R w KQ -",            # eligibility with the figurehead off its home square (a figurehead is conjured on c1 / g1)
    "4k3/8/8/8/8/NNN5/PPPPPPPP/RNBQK3 w K -",  # sixteen figurines, east eligibility, no cop on h1: secret service makes a seventeenth
    "r3k3/pppppppp/nnnnnn2/8/8/8/8/4K3 b kq -",  # the same for Blue (sixteen, the h8 corner empty)
]


def outside_w_worlds():
    return np.array([O.reconstruct(f) for f in OUTSIDE_W], dtype=O.WORLD_DTYPE)


def is_well_formed(w):
    """Domain W of include/leafline_b200.h, restated on the host for the closure test."""
    planes = [int(x) for x in w["plane"]]
    acc = 0
    for p in planes:
        if acc & p:
            return False
        acc |= p
    if bin(planes[5]).count("1") > 1 or bin(planes[11]).count("1") > 1:
        return False
    orange, blue = 0, 0
    for j in range(6):
        orange, blue = orange | planes[j], blue | planes[6 + j]
    if bin(orange).count("1") > 16 or bin(blue).count("1") > 16:
        return False
    e = int(w["service_eligibility"])
    if e & 3 and planes[5] & ~(1 << 4):
        return False
    if e & 12 and planes[11] & ~(1 << 60):
        return False
    # a set eligibility bit wants the cop on its corner or a spare seat among the sixteen (secret service conjures a cop)
    if bin(orange).count("1") == 16 and ((e & 1 and not planes[3] & 1) or (e & 2 and not planes[3] & (1 << 7))):
        return False
    if bin(blue).count("1") == 16 and ((e & 4 and not planes[9] & (1 << 56)) or (e & 8 and not planes[9] & (1 << 63))):
        return False
    pb = int(w["passing_by"])
    if pb != 0xFF:
        r, f = pb >> 4, pb & 15
        if r > 7 or f > 7:
            return False
        bit = 1 << (8 * r + f)
        if acc & bit:
            return False
        if int(w["initiative"]) == 0:
            if r != 5 or not planes[6] & (bit >> 8):
                return False
        else:
            if r != 2 or not planes[0] & (bit << 8):
                return False
    return True
