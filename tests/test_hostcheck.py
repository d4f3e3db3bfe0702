"""Host-side logic check of the move generator the CUDA kernels are built from.

tests/hostcheck/hostcheck.cpp compiles leafline_b200/csrc/ll_device.cuh as plain C++ (test
infrastructure only -- the shipped library has no CPU path) so the legality-by-masks generator, the
set-wise move counter, make-move and the scorer can be compared with the oracle over tens of
thousands of positions without a GPU.  The GPU parity tests (test_gpu_parity.py) then check the very
same code compiled for sm_100a."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from positions import named_worlds, reckless_walk_worlds

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
SO = os.path.join(HERE, "hostcheck", "_hostcheck.so")
HDR = os.path.join(HERE, "..", "leafline_b200", "csrc", "ll_device.cuh")


@pytest.fixture(scope="module")
def hc():
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-o", SO, SRC], check=True)
    lib = C.CDLL(SO)
    lib.hc_count.argtypes = [C.c_void_p, C.c_int]
    lib.hc_lookahead.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.hc_score.argtypes = [C.c_void_p]
    lib.hc_score.restype = C.c_float
    lib.hc_well_formed.argtypes = [C.c_void_p]
    lib.hc_lookahead_literal.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.hc_domain_class.argtypes = [C.c_void_p]
    lib.hc_record_moves.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.hc_generator_moves.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    lib.hc_legal_record_moves.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    lib.hc_legal_generator_moves.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    lib.hc_feature_check.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.hc_canonical_key.argtypes = [C.c_uint32]
    lib.hc_quiet_ways_check.argtypes = [C.c_void_p, C.POINTER(C.c_int)]
    lib.hc_canonical_key.restype = C.c_uint32
    lib.hc_init()
    return lib


def quiet_commit(w, d):
    """nobody hospitalized, nothing ascended, no secret service (a passing-by hospitalizes): what k_horizon's quiet path takes"""
    service = int(d["job"]) == 5 and abs((int(d["whence"]) & 15) - (int(d["whither"]) & 15)) == 2
    return int(d["hospitalization"]) == O.NONE and int(d["ascension"]) == O.NONE and not service


def check(hc, worlds):
    buf = np.zeros(1024, dtype=O.COMMIT_DTYPE)
    for w in worlds:
        w = np.ascontiguousarray(w)
        ptr = w.ctypes.data_as(C.c_void_p)
        assert hc.hc_well_formed(ptr) == 1
        for nihil, ref in ((0, O.lookahead), (1, O.reckless_lookahead)):
            want = ref(w)
            n = hc.hc_lookahead(ptr, nihil, buf.ctypes.data_as(C.c_void_p), 1024)
            assert n == len(want), (O.preserve(w), nihil, n, len(want))
            assert buf[:n].tobytes() == want.tobytes(), (O.preserve(w), nihil)
            assert hc.hc_count(ptr, nihil) == n, (O.preserve(w), nihil)
        assert np.float32(hc.hc_score(ptr)).tobytes() == np.float32(O.score(w)).tobytes()
        # the horizon kernel's evaluation by features (+-1 updates of the parent's integers) equals score(make_child(...))
        checked = C.c_int(0)
        assert hc.hc_feature_check(ptr, C.byref(checked)) == 0, O.preserve(w)
        assert checked.value == len(want)
        # ... and by the ways a record's quiet commits change the features, set-wise
        quiet = C.c_int(0)
        assert hc.hc_quiet_ways_check(ptr, C.byref(quiet)) == 0, O.preserve(w)
        assert quiet.value == sum(1 for d in want if quiet_commit(w, d)), O.preserve(w)
        # the record form of the reckless commits (horizon kernel) holds exactly the generator's commits, in another order
        a, b, nrec = np.zeros(1024, dtype=np.uint32), np.zeros(1024, dtype=np.uint32), C.c_int(0)
        for stuns in (0, 1):
            na = hc.hc_record_moves(ptr, stuns, a.ctypes.data_as(C.c_void_p), 1024, C.byref(nrec))
            nb = hc.hc_generator_moves(ptr, stuns, b.ctypes.data_as(C.c_void_p), 1024)
            assert 0 <= nrec.value <= 36 and na == nb, (O.preserve(w), stuns, na, nb, nrec.value)
            assert sorted(a[:na]) == sorted(b[:nb]), (O.preserve(w), stuns)
        # ... and the legal records (perft chain / tail) hold exactly lookahead()'s commits
        na = hc.hc_legal_record_moves(ptr, a.ctypes.data_as(C.c_void_p), 1024, C.byref(nrec))
        nb = hc.hc_legal_generator_moves(ptr, b.ctypes.data_as(C.c_void_p), 1024)
        assert na == nb and sorted(a[:na]) == sorted(b[:nb]), (O.preserve(w), na, nb)
        # ... which the canonical key sorts back into the reference's order (the root ply of the perft chain does that)
        keys = [hc.hc_canonical_key(int(m)) for m in b[:nb]]
        assert keys == sorted(keys) and len(set(keys)) == nb, O.preserve(w)
        for stuns in (0,):
            nr = hc.hc_generator_moves(ptr, 0, b.ctypes.data_as(C.c_void_p), 1024)
            keys = [hc.hc_canonical_key(int(m)) for m in b[:nr]]
            assert keys == sorted(keys) and len(set(keys)) == nr, O.preserve(w)


def test_named_positions(hc):
    check(hc, named_worlds())


def test_playout_positions(hc):
    check(hc, O.playout_batch(0x1EAF11E, 0, 3000))


def test_reckless_walk_positions(hc):
    check(hc, reckless_walk_worlds(seed=23, games=60, plies=70))


def test_quiet_ways_over_many_positions(hc):
    """The horizon kernel's set-wise valuation of quiet commits, held against score(make_child(...)) commit by commit and against
    the commit-by-commit ways record by record (hc_feature_check, hc_quiet_ways_check) on a hundred thousand positions:
    pseudo-legal walks (captured figureheads, phantom castling) and arbitrary members of the domain.  (A one-off sweep over
    200 000 playout positions more, 8.9 M quiet commits in all, found nothing either.)"""
    seen = 0
    for worlds in (reckless_walk_worlds(seed=5, games=1000, plies=90), random_domain_worlds(31, 20000)):
        for w in worlds:
            w = np.ascontiguousarray(w)
            ptr = w.ctypes.data_as(C.c_void_p)
            checked, quiet = C.c_int(0), C.c_int(0)
            assert hc.hc_feature_check(ptr, C.byref(checked)) == 0, O.preserve(w)
            assert hc.hc_quiet_ways_check(ptr, C.byref(quiet)) == 0, O.preserve(w)
            seen += quiet.value
    assert seen > 2_000_000


def random_domain_worlds(seed, n):
    """Arbitrary members of the ABI's domain W (include/leafline_b200.h): up to 16 figurines a team anywhere on the board --
    servants on first and last ranks, a dozen princesses, no figurehead at all -- with a consistent eligibility."""
    from positions import is_well_formed

    rng = np.random.default_rng(seed)
    out = []
    while len(out) < n:
        w = np.zeros((), dtype=O.WORLD_DTYPE)
        squares = rng.permutation(64)
        k = 0
        for team in range(2):
            count = int(rng.integers(1, 17))
            has_figurehead = rng.random() < 0.85
            for i in range(count):
                sq = int(squares[k])
                k += 1
                job = 5 if (i == 0 and has_figurehead) else int(rng.choice([0, 0, 0, 1, 2, 3, 4]))
                w["plane"][6 * team + job] |= np.uint64(1) << np.uint64(sq)
        w["initiative"] = int(rng.integers(0, 2))
        elig = 0
        if int(w["plane"][5]) in (0, 1 << 4):
            elig |= int(rng.integers(0, 4))
        if int(w["plane"][11]) in (0, 1 << 60):
            elig |= int(rng.integers(0, 4)) << 2
        w["service_eligibility"] = elig
        w["passing_by"] = 0xFF
        if is_well_formed(w):  # (a full team with an eligibility bit and no cop on that corner belongs to the literal domain)
            out.append(w)
    return np.array(out, dtype=O.WORLD_DTYPE)


def test_arbitrary_positions_of_the_domain(hc):
    # the ABI accepts every well-formed position, not only reachable ones: generator, counts, records, features and
    # score must follow the reference's rules there too
    check(hc, random_domain_worlds(20261019, 1500))


def test_single_escape_square_experiment_keeps_parity(hc):
    # ll_device.cuh carries one experiment that is off in the product build (LL_SINGLE_ESCAPE_TEST: a figurehead with one
    # escape square asks whether that square is attacked instead of filling the danger map).  Built here with the switch
    # on, it must still equal the oracle everywhere, so that it can be timed on the GPU at any moment.
    so = SO.replace("_hostcheck.so", "_hostcheck_single_escape.so")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(SRC), os.path.getmtime(HDR)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-w", "-DLL_SINGLE_ESCAPE_TEST", "-o", so, SRC], check=True)
    lib = C.CDLL(so)
    for name in ("hc_count", "hc_lookahead", "hc_score", "hc_well_formed", "hc_record_moves", "hc_generator_moves", "hc_legal_record_moves",
                 "hc_legal_generator_moves", "hc_feature_check", "hc_canonical_key"):
        getattr(lib, name).argtypes = getattr(hc, name).argtypes
        getattr(lib, name).restype = getattr(hc, name).restype
    lib.hc_init()
    check(lib, named_worlds())
    check(lib, O.playout_batch(0x1EAF11E, 5000, 1200))
    check(lib, reckless_walk_worlds(seed=41, games=25, plies=60))
    check(lib, random_domain_worlds(7, 600))


def test_checks_pins_and_passing_by_corners(hc):
    fens = [
        "4k3/8/8/8/8/8/4r3/4K3 w - -",          # adjacent checker, capture by the figurehead only
        "4k3/8/8/8/7b/8/5P2/4K3 w - -",         # servant pinned on a diagonal
        "4k3/4r3/8/8/8/8/4P3/4K3 w - -",        # servant pinned on a file: may still push
        "4k3/8/8/8/1b6/8/3P4/4K3 w - -",        # pinned servant, nothing to stun
        "4k3/8/8/8/1b6/2p5/3P4/4K3 w - -",      # pinned servant may stun along the pin
        "8/8/8/8/k2Pp2R/8/8/4K3 b - d3",        # passing by would uncover the cop on the rank
        "8/8/8/2k5/3Pp3/8/8/4K3 b - d3",        # the double-stepped servant gives check: passing by answers it
        "4k3/8/8/8/8/2n5/8/R3K2R w KQ -",       # pony check: no secret service, block impossible
        "4k3/8/8/8/8/8/3n1n2/R3K2R w KQ -",     # double check
        "r3k2r/8/8/8/8/8/8/R3K2R w KQkq -", "r3k2r/8/8/8/8/8/8/R3K2R b KQkq -",
        "4k3/8/8/8/8/8/8/RN2K2R w KQ -", "4k2r/8/8/8/8/8/8/4K2R w Kk -",
        "3rk3/8/8/8/8/8/8/R3K2R w KQ -",        # d1 leered at: West refused, East fine
        "1r2k3/8/8/8/8/8/8/R3K2R w KQ -",       # b1 leered at: West refused (quirk 1)
        "4k3/1P6/8/8/8/8/8/4K3 w - -", "n1n1k3/1P6/8/8/8/8/8/4K3 w - -", "4k3/8/8/8/8/8/1p6/N1N1K3 b - -",
        "4k3/8/8/8/8/8/8/4K2n w K -", "4k3/8/8/8/8/8/8/7R w K -", "4k3/8/8/8/8/8/8/4r2R w K -",
        "Q3k3/8/8/8/8/8/8/q3K3 w - -", "7k/6Q1/8/8/8/8/8/K7 b - -", "7k/5KQ1/8/8/8/8/8/8 b - -",
        "8/8/8/8/8/1k6/8/K7 w - -", "8/8/8/8/8/8/1k6/K7 w - -",  # adjacent figureheads (reckless lines get there)
        "k7/8/8/3q4/8/8/3R4/3K4 w - -", "k7/8/8/3q4/8/3B4/8/3K4 w - -", "k7/8/8/q7/8/2Q5/8/4K3 w - -",
    ]
    worlds = [O.reconstruct(f) for f in fens]
    check(hc, worlds)
    # and everything one and two plies below them (reckless: figureheads get stunned, phantom service)
    deeper = []
    for w in worlds:
        for c in O.reckless_lookahead(w):
            deeper.append(c["tree"].copy())
            for d in O.reckless_lookahead(c["tree"])[:12]:
                deeper.append(d["tree"].copy())
    check(hc, deeper)


def walk(worlds, plies, seed):
    """the given positions and seeded reckless / legal descendants of them"""
    import random

    rng = random.Random(seed)
    out = []
    for w in worlds:
        for reckless in (False, True):
            cur = w.copy()
            for _ in range(plies):
                out.append(cur.copy())
                kids = O.reckless_lookahead(cur) if reckless else O.lookahead(cur)
                if not len(kids):
                    break
                cur = kids[rng.randrange(len(kids))]["tree"].copy()
    return out


def test_literal_path_answers_positions_outside_domain_w(hc):
    """Two figureheads a team, seventeen figurines, eligibility without figurehead or cop in place (life.rs:678-699 loops over any
    number of figureheads): the literal generator's commits equal the oracle's, byte for byte and in order, on such positions and
    on seeded descendants of them; on positions inside W it equals the fast generator."""
    from positions import is_well_formed, outside_w_worlds

    roots = outside_w_worlds()
    assert all(hc.hc_domain_class(np.ascontiguousarray(w).ctypes.data_as(C.c_void_p)) == 1 for w in roots)
    worlds = walk(roots, 12, seed=3) + list(named_worlds()) + list(reckless_walk_worlds(seed=9, games=6, plies=40))
    buf = np.zeros(1024, dtype=O.COMMIT_DTYPE)
    outside = 0
    for w in worlds:
        w = np.ascontiguousarray(w).reshape(())
        ptr = w.ctypes.data_as(C.c_void_p)
        cls = hc.hc_domain_class(ptr)
        assert cls in (0, 1) and (cls == 0) == is_well_formed(w), O.preserve(w)
        outside += cls
        for nihil, ref in ((0, O.lookahead), (1, O.reckless_lookahead)):
            want = ref(w)
            n = hc.hc_lookahead_literal(ptr, nihil, buf.ctypes.data_as(C.c_void_p), 1024)
            assert n == len(want) and buf[:n].tobytes() == want.tobytes(), (O.preserve(w), nihil, n, len(want))
    assert outside > 100
