"""leafline_b200 -- B200-native leaf layer (legal move generation + static scoring) for Leafline.

Host-side mirror of the reference's interface for this path, over the C ABI in
include/leafline_b200.h (hand-written sm_100a CUDA, leafline_b200/csrc/):

    reference (Rust)                                  here
    ------------------------------------------------  ------------------------------------------
    WorldState::new()              life.rs:184-221    new_world()
    WorldState::reconstruct(s)     life.rs:428-507    reconstruct(s)
    world.preserve()               life.rs:293-360    preserve(world)
    world.lookahead()              life.rs:1078       Engine.lookahead(worlds)
    world.reckless_lookahead()     life.rs:1082       Engine.reckless_lookahead(worlds)
    world.in_critical_endangerment life.rs:678        Engine.in_critical_endangerment(worlds, teams)
    mind::score(world)             mind.rs:50         Engine.score(worlds)
    α_β_negamax_search leaf rule   mind.rs:279-287    Engine.horizon(worlds, plies)
    mind::kickoff(world, depth..)  mind.rs:441        Engine.kickoff(world, depth, nihilistically)
    kickoff's thread-per-root-move fan-out mind.rs:399-437  MultiEngine.kickoff / .perft (several GPUs, one tree)

Positions are numpy records of WORLD_DTYPE (= ll_world_t, 104 bytes), commits of COMMIT_DTYPE
(= ll_commit_t).  Everything is batch-first.  There is no CPU path: constructing an Engine
without the built library or without a CUDA device raises.
"""
import ctypes as C

import numpy as np

from . import _abi
from ._abi import COMMIT_DTYPE, FORECAST_DTYPE, WORLD_DTYPE, LeaflineError, Soa  # noqa: F401
from .runes import preserve, reconstruct  # noqa: F401

ORANGE, BLUE = 0, 1
SERVANT, PONY, SCHOLAR, COP, PRINCESS, FIGUREHEAD = range(6)
NONE = 0xFF
PLAYOUT_SEED = 0x1EAF11E
PERFT_DEVICE_WORDS = 8 + 1024


def locale(rank, file):
    """space.rs:34 packing of a Locale."""
    return (rank << 4) | file


def algebraic(name):
    """space.rs:53-63."""
    return locale(ord(name[1]) - 49, ord(name[0]) - 97)


def to_algebraic(loc):
    """space.rs:49-51."""
    return "abcdefgh"[loc & 15] + str((loc >> 4) + 1)


def agent(team, job):
    return (team << 3) | job


def pagan_variation(forecast):
    """The variation of one forecast as "e2e4 e7e5 ..." (cf. pagan_variation_format, mind.rs:174-180)."""
    n = int(forecast["variation_len"])
    return " ".join(to_algebraic(p["whence"]) + to_algebraic(p["whither"]) for p in forecast["variation"][:n])


def new_world():
    """WorldState::new() (life.rs:184-221)."""
    return reconstruct("rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -")


def _worlds(worlds):
    a = np.ascontiguousarray(worlds, dtype=WORLD_DTYPE)
    return a.reshape(-1)


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


class Engine:
    """One ll_ctx: a CUDA device, a stream and the HBM scratch of the leaf layer.  Not thread-safe."""

    def __init__(self, device=0, _borrowed_ctx=None):
        self._lib = _abi.load()
        self._ctx = C.c_void_p()
        self._owned = _borrowed_ctx is None
        if self._owned:
            _abi.check(self._lib.ll_init(device, C.byref(self._ctx)))
        else:
            self._ctx = C.c_void_p(_borrowed_ctx)  # a member of an ll_mctx: the group shuts it down
        self.device = device

    def close(self):
        if self._ctx and self._owned:
            self._lib.ll_shutdown(self._ctx)
        self._ctx = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- plumbing
    def set_stream(self, cuda_stream):
        _abi.check(self._lib.ll_set_stream(self._ctx, C.c_void_p(cuda_stream)))

    def synchronize(self):
        _abi.check(self._lib.ll_synchronize(self._ctx))

    @property
    def sm_count(self):
        return self._lib.ll_device_sm_count(self._ctx)

    @property
    def launch_count(self):
        return int(self._lib.ll_launch_count(self._ctx))

    def timing(self, on=True):
        _abi.check(self._lib.ll_timing_enable(self._ctx, int(on)))

    def timing_reset(self):
        _abi.check(self._lib.ll_timing_reset(self._ctx))

    def timing_query(self, kernel):
        ms, launches = C.c_double(0), C.c_uint64(0)
        _abi.check(self._lib.ll_timing_query(self._ctx, kernel.encode(), C.byref(ms), C.byref(launches)))
        return ms.value, launches.value

    def force_synced(self, on=True):
        """Testing aid: True/1 = level-by-level perft, 2 = the launch chain without its CUDA graph, False/0 = default."""
        _abi.check(self._lib.ll_debug_force_synced(self._ctx, int(on)))

    def selfcheck_soa(self, soa):
        """Testing aid -> (disagreeing positions, first index or None, legal commits, reckless commits); see the header."""
        out = np.zeros(4, dtype=np.uint64)
        _abi.check(self._lib.ll_debug_selfcheck_soa(self._ctx, C.byref(soa), _ptr(out)))
        return int(out[0]), (None if out[1] == np.uint64(0xFFFFFFFFFFFFFFFF) else int(out[1])), int(out[2]), int(out[3])

    def debug_tables(self):
        """(pony, figurehead) movement tables as resident on the device, 64 u64 each"""
        pony, fig = np.zeros(64, dtype=np.uint64), np.zeros(64, dtype=np.uint64)
        _abi.check(self._lib.ll_debug_tables(self._ctx, _ptr(pony), _ptr(fig)))
        return pony, fig

    def debug_horizon_overflowed(self):
        """Testing aid: parents of the last level-by-level horizon call that went to the overflow kernel; see the header."""
        out = np.zeros(1, dtype=np.uint32)
        _abi.check(self._lib.ll_debug_horizon_overflowed(self._ctx, _ptr(out)))
        return int(out[0])

    def measure_int_peak(self, kind=0):
        """Sustained thread-level integer instructions/s (0: LOP3 stream, 1: LOP3+IMAD) for the integer roofline."""
        v = C.c_double(0)
        _abi.check(self._lib.ll_measure_int_peak(self._ctx, kind, C.byref(v)))
        return v.value

    # ---- the leaf layer on host buffers
    def score(self, worlds):
        """mind::score for every world (mind.rs:50-143) -> float32[n]."""
        w = _worlds(worlds)
        out = np.empty(len(w), dtype=np.float32)
        _abi.check(self._lib.ll_score_batch(self._ctx, _ptr(w), len(w), _ptr(out)))
        return out

    def score_into(self, worlds, out):
        """score() into a caller-owned float32 host buffer (e.g. pinned memory): no allocation, no copy of the input."""
        w = _worlds(worlds)
        assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"] and len(out) >= len(w)
        _abi.check(self._lib.ll_score_batch(self._ctx, _ptr(w), len(w), _ptr(out)))
        return out

    def lookahead_counts(self, worlds, nihilistically=False):
        w = _worlds(worlds)
        offsets = np.zeros(len(w) + 1, dtype=np.uint32)
        _abi.check(self._lib.ll_lookahead_batch(self._ctx, _ptr(w), len(w), int(nihilistically), _ptr(offsets), None, 0))
        return np.diff(offsets.astype(np.int64))

    def lookahead(self, worlds, nihilistically=False):
        """lookahead() / reckless_lookahead() of every world -> (offsets[n+1], commits) in reference order."""
        w = _worlds(worlds)
        offsets = np.zeros(len(w) + 1, dtype=np.uint32)
        _abi.check(self._lib.ll_lookahead_batch(self._ctx, _ptr(w), len(w), int(nihilistically), _ptr(offsets), None, 0))
        total = int(offsets[-1])
        commits = np.zeros(total, dtype=COMMIT_DTYPE)
        if total:
            _abi.check(
                self._lib.ll_lookahead_batch(self._ctx, _ptr(w), len(w), int(nihilistically), _ptr(offsets), _ptr(commits), total)
            )
        return offsets, commits

    def reckless_lookahead(self, worlds):
        return self.lookahead(worlds, nihilistically=True)

    def in_critical_endangerment(self, worlds, teams):
        w = _worlds(worlds)
        t = np.ascontiguousarray(np.broadcast_to(np.asarray(teams, dtype=np.uint8), (len(w),)))
        out = np.zeros(len(w), dtype=np.uint8)
        _abi.check(self._lib.ll_in_critical_endangerment_batch(self._ctx, _ptr(w), _ptr(t), len(w), _ptr(out)))
        return out.astype(bool)

    def perft(self, world, depth, shard=(0, 1)):
        """Leaf count of the legal-move tree -> (leaves, per_root[n_roots]) for this shard."""
        w = _worlds(world)
        leaves, n_roots = C.c_uint64(0), C.c_uint32(0)
        per_root = np.zeros(1024, dtype=np.uint64)
        _abi.check(
            self._lib.ll_perft(self._ctx, _ptr(w), depth, shard[0], shard[1], C.byref(leaves), _ptr(per_root), 1024, C.byref(n_roots))
        )
        return int(leaves.value), per_root[: n_roots.value].copy()

    def perft_device(self, world, depth, d_out_ptr, shard=(0, 1)):
        """Asynchronous perft whose counts stay in DEVICE memory (u64[PERFT_DEVICE_WORDS] at d_out_ptr:
        [0] leaves, [1] root moves, [2] overflow flag, [8+i] per-root) -- for an on-device all-reduce."""
        w = _worlds(world)
        _abi.check(self._lib.ll_perft_device(self._ctx, _ptr(w), depth, shard[0], shard[1], C.c_void_p(d_out_ptr)))

    def horizon(self, worlds, plies, quiet=None):
        """Exact negamax value of each frontier node over `plies` reckless plies (mind.rs:279-301);
        quiet = None or the quiescence extension (mind.rs:284-300)."""
        w = _worlds(worlds)
        out = np.empty(len(w), dtype=np.float32)
        _abi.check(self._lib.ll_horizon_batch(self._ctx, _ptr(w), len(w), plies, -1 if quiet is None else int(quiet), _ptr(out)))
        return out

    def kickoff(self, world, depth, nihilistically=False, shard=(0, 1), quiet=None):
        """mind::kickoff without TT/history: (root commits, root scores, leaves scored), canonical order."""
        w = _worlds(world)
        roots = np.zeros(1024, dtype=COMMIT_DTYPE)
        scores = np.zeros(1024, dtype=np.float32)
        n_roots, scored = C.c_uint32(0), C.c_uint64(0)
        _abi.check(
            self._lib.ll_kickoff(
                self._ctx, _ptr(w), depth, -1 if quiet is None else int(quiet), int(nihilistically), shard[0], shard[1], _ptr(roots), _ptr(scores), 1024,
                C.byref(n_roots), C.byref(scored),
            )
        )
        return roots[: n_roots.value].copy(), scores[: n_roots.value].copy(), int(scored.value)

    # ---- the kickoff family as the reference's callers see it (mind.rs:441-490)
    def forecast(self, world, depth, nihilistically=False, quiet=None):
        """kickoff::<Variation>: FORECAST_DTYPE records (commit, score, variation) sorted by score, best first."""
        w = _worlds(world)
        out = np.zeros(1024, dtype=FORECAST_DTYPE)
        n, scored = C.c_uint32(0), C.c_uint64(0)
        _abi.check(self._lib.ll_forecast(self._ctx, _ptr(w), depth, -1 if quiet is None else int(quiet), int(nihilistically),
                                         _ptr(out), 1024, C.byref(n), C.byref(scored)))
        return out[: n.value].copy()

    def deep_forecast(self, world, depth, nihilistically=False, quiet=None):
        """kickoff to any depth: full-width on the device (with variations) while the tree fits HBM, the host alpha-beta
        driver over device frontier batches beyond that (same scores; the variation is the root move)."""
        if depth < 8:
            try:
                return self.forecast(world, depth, nihilistically=nihilistically, quiet=quiet)
            except LeaflineError as e:
                if e.code != _abi.ERR_OOM:
                    raise
        return self.alpha_beta_forecast(world, depth, device_plies=5 if depth >= 9 else 4, nihilistically=nihilistically, quiet=quiet)[0]

    def alpha_beta_forecast(self, world, depth, device_plies=4, nihilistically=False, quiet=None):
        """The reference's search structure: alpha-beta with MVV-LVA + history ordering on the host, frontier nodes with
        `device_plies` plies left valued exactly on the device -> (forecasts, host nodes expanded, frontier nodes handed off);
        `last_search_stats` also holds how many values the (sound) memory bank supplied."""
        w = _worlds(world)
        out = np.zeros(1024, dtype=FORECAST_DTYPE)
        n = C.c_uint32(0)
        stats = np.zeros(3, dtype=np.uint64)
        _abi.check(self._lib.ll_alpha_beta_forecast(self._ctx, _ptr(w), depth, device_plies, -1 if quiet is None else int(quiet),
                                                    int(nihilistically), _ptr(out), 1024, C.byref(n), _ptr(stats)))
        self.last_search_stats = {"host_nodes_expanded": int(stats[0]), "frontier_nodes_handed_to_device": int(stats[1]),
                                  "values_remembered": int(stats[2])}
        return out[: n.value].copy(), int(stats[0]), int(stats[1])

    def set_search_threads(self, threads):
        """Searcher threads of the host alpha-beta driver (one per root move up to this many, mind.rs:399-414)."""
        _abi.check(self._lib.ll_set_search_threads(self._ctx, int(threads)))

    def set_deja_vu_bound(self, gib):
        """The `déjà_vu_bound` of the reference's kickoff functions: GiB of memory bank for this engine's searches."""
        _abi.check(self._lib.ll_set_deja_vu_bound(self._ctx, float(gib)))

    def fixed_depth_sequence_kickoff(self, world, depth_sequence, nihilistically=False):
        """mind.rs:473-490."""
        w = _worlds(world)
        depths = np.ascontiguousarray(depth_sequence, dtype=np.uint8)
        out = np.zeros(1024, dtype=FORECAST_DTYPE)
        n = C.c_uint32(0)
        _abi.check(self._lib.ll_fixed_depth_sequence_forecast(self._ctx, _ptr(w), _ptr(depths), len(depths), int(nihilistically),
                                                              _ptr(out), 1024, C.byref(n)))
        return out[: n.value].copy()

    def iterative_deepening_kickoff(self, world, timeout_seconds, nihilistically=False):
        """mind.rs:451-469 -> (forecasts of the deepest iteration that finished in time, its depth)."""
        w = _worlds(world)
        out = np.zeros(1024, dtype=FORECAST_DTYPE)
        n, depth = C.c_uint32(0), C.c_uint32(0)
        _abi.check(self._lib.ll_iterative_deepening_forecast(self._ctx, _ptr(w), float(timeout_seconds), int(nihilistically),
                                                             _ptr(out), 1024, C.byref(n), C.byref(depth)))
        return out[: n.value].copy(), int(depth.value)

    # ---- positions resident in HBM (SoA planes)
    def soa_alloc(self, n):
        soa = Soa()
        _abi.check(self._lib.ll_soa_alloc(self._ctx, n, C.byref(soa)))
        return soa

    def soa_free(self, soa):
        _abi.check(self._lib.ll_soa_free(self._ctx, C.byref(soa)))

    def soa_upload(self, worlds, soa=None):
        w = _worlds(worlds)
        if soa is None:
            soa = self.soa_alloc(len(w))
        _abi.check(self._lib.ll_soa_upload(self._ctx, _ptr(w), len(w), C.byref(soa)))
        return soa

    def soa_download(self, soa, first=0, n=None):
        n = soa.n - first if n is None else n
        out = np.zeros(n, dtype=WORLD_DTYPE)
        _abi.check(self._lib.ll_soa_download(self._ctx, C.byref(soa), first, n, _ptr(out)))
        return out

    def score_soa(self, soa, d_scores_ptr):
        """Asynchronous: scores of resident positions into the DEVICE buffer at d_scores_ptr."""
        _abi.check(self._lib.ll_score_soa(self._ctx, C.byref(soa), C.c_void_p(d_scores_ptr)))

    def horizon_soa(self, soa, plies, d_values_ptr, quiet=None):
        scored = C.c_uint64(0)
        _abi.check(self._lib.ll_horizon_soa(self._ctx, C.byref(soa), plies, -1 if quiet is None else int(quiet), C.c_void_p(d_values_ptr), C.byref(scored)))
        return int(scored.value)

    def playout_soa(self, n, seed=PLAYOUT_SEED, first_index=0, soa=None):
        """Seeded random-playout positions generated on the device (DESIGN.md "playout set")."""
        if soa is None:
            soa = self.soa_alloc(n)
        _abi.check(self._lib.ll_playout_soa(self._ctx, seed, first_index, n, C.byref(soa)))
        return soa


class MultiEngine:
    """One ll_mctx: several GPUs sharing ONE tree behind the C ABI (the reference's fan-out inside the engine,
    mind.rs:399-437).  The group owns a context per GPU, the NCCL communicator and the device-resident reduction.

        MultiEngine(devices=[0, 1, ...])                      one process, one host thread + stream per device
        MultiEngine.for_rank(device, rank, world, unique_id)  one process per GPU (torchrun): every call is collective
    """

    def __init__(self, devices=None, _handle=None):
        self._lib = _abi.load()
        self._g = C.c_void_p()
        if _handle is not None:
            self._g = _handle
        else:
            arr = (C.c_int * len(devices))(*devices)
            _abi.check(self._lib.ll_init_multi(arr, len(devices), C.byref(self._g)))

    @staticmethod
    def unique_id():
        """128 bytes naming a new communicator; carry them to every process (ncclGetUniqueId)."""
        buf = (C.c_ubyte * 128)()
        _abi.check(_abi.load().ll_multi_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    @classmethod
    def for_rank(cls, device, rank, world, unique_id):
        lib = _abi.load()
        g = C.c_void_p()
        buf = (C.c_ubyte * 128)(*unique_id) if unique_id is not None else None
        _abi.check(lib.ll_init_multi_rank(device, rank, world, C.cast(buf, C.c_void_p) if buf is not None else None, C.byref(g)))
        return cls(_handle=g)

    def close(self):
        if self._g:
            self._lib.ll_shutdown_multi(self._g)
            self._g = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def world(self):
        return self._lib.ll_multi_world(self._g)

    @property
    def local_count(self):
        return self._lib.ll_multi_local_count(self._g)

    @property
    def reduce_path(self):
        return "peer-memory" if self._lib.ll_multi_reduce_path(self._g) else "nccl"

    def member(self, local_index=0):
        """The member's ll_ctx as an Engine (borrowed: set_stream, timing, launch_count ...)."""
        ctx = self._lib.ll_multi_ctx(self._g, local_index)
        if not ctx:
            raise IndexError(local_index)
        return Engine(_borrowed_ctx=ctx)

    def set_shard_min_depth(self, perft_depth=0, kickoff_depth=0):
        _abi.check(self._lib.ll_multi_set_shard_min_depth(self._g, perft_depth, kickoff_depth))

    def perft(self, world, depth):
        """ll_perft_multi -> (leaves, per_root[n_roots]) of the whole tree."""
        w = _worlds(world)
        leaves, n_roots = C.c_uint64(0), C.c_uint32(0)
        per_root = np.zeros(1024, dtype=np.uint64)
        _abi.check(self._lib.ll_perft_multi(self._g, _ptr(w), depth, C.byref(leaves), _ptr(per_root), 1024, C.byref(n_roots)))
        return int(leaves.value), per_root[: n_roots.value].copy()

    def perft_batch(self, worlds, depth, out=None):
        """ll_perft_batch_multi: leaf counts of independent roots, root i counted by rank i % world -> uint64[n]."""
        w = _worlds(worlds)
        if out is None:
            out = np.zeros(len(w), dtype=np.uint64)
        _abi.check(self._lib.ll_perft_batch_multi(self._g, _ptr(w), len(w), depth, _ptr(out)))
        return out

    def kickoff(self, world, depth, nihilistically=False, quiet=None):
        """ll_kickoff_multi -> (root commits, root scores, leaves scored), canonical order, on every rank."""
        w = _worlds(world)
        roots = np.zeros(1024, dtype=COMMIT_DTYPE)
        scores = np.zeros(1024, dtype=np.float32)
        n_roots, scored = C.c_uint32(0), C.c_uint64(0)
        _abi.check(self._lib.ll_kickoff_multi(self._g, _ptr(w), depth, -1 if quiet is None else int(quiet), int(nihilistically),
                                              _ptr(roots), _ptr(scores), 1024, C.byref(n_roots), C.byref(scored)))
        return roots[: n_roots.value].copy(), scores[: n_roots.value].copy(), int(scored.value)

    def alpha_beta_forecast(self, world, depth, device_plies=4, nihilistically=False, quiet=None):
        """ll_alpha_beta_forecast_multi: root moves searched concurrently, one subset per GPU -> (forecasts, host nodes
        expanded, frontier nodes handed off), summed over the ranks."""
        w = _worlds(world)
        out = np.zeros(1024, dtype=FORECAST_DTYPE)
        n = C.c_uint32(0)
        stats = np.zeros(3, dtype=np.uint64)
        _abi.check(self._lib.ll_alpha_beta_forecast_multi(self._g, _ptr(w), depth, device_plies, -1 if quiet is None else int(quiet),
                                                          int(nihilistically), _ptr(out), 1024, C.byref(n), _ptr(stats)))
        self.last_search_stats = {"host_nodes_expanded": int(stats[0]), "frontier_nodes_handed_to_device": int(stats[1]),
                                  "values_remembered": int(stats[2])}
        return out[: n.value].copy(), int(stats[0]), int(stats[1])
