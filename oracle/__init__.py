"""ctypes binding of the CPU oracle (oracle/leafline_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (leafline_b200/) never imports this.
Parity pinning: see leafline_oracle.h (the Rust reference cannot be built here; the restatement is
pinned by the reference's own known-answer tests, tests/test_oracle_golden.py).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libleafline_oracle.so")

ORANGE, BLUE = 0, 1
SERVANT, PONY, SCHOLAR, COP, PRINCESS, FIGUREHEAD = range(6)
NONE = 0xFF
MAX_COMMITS = 512

WORLD_DTYPE = np.dtype(
    [("plane", "<u8", (12,)), ("initiative", "u1"), ("service_eligibility", "u1"), ("passing_by", "u1"),
     ("pad", "u1", (5,))]
)
COMMIT_DTYPE = np.dtype(
    [("team", "u1"), ("job", "u1"), ("whence", "u1"), ("whither", "u1"), ("hospitalization", "u1"),
     ("ascension", "u1"), ("pad", "u1", (2,)), ("tree", WORLD_DTYPE)]
)
assert WORLD_DTYPE.itemsize == 104 and COMMIT_DTYPE.itemsize == 112


def _host_signature():
    """what `-march=native` means on this machine: the library is rebuilt when it was compiled for another CPU
    (the built .so travels to the GPU box with the snapshot; its host cores are not this container's)"""
    flags = ""
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("flags"):
                    flags = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    import hashlib

    return hashlib.sha256(flags.encode()).hexdigest()[:16]


def build(force=False):
    """Compile the oracle with gcc (oracle/Makefile: -O3 -march=native -ffp-contract=off)."""
    src = [os.path.join(_HERE, f) for f in ("leafline_oracle.c", "leafline_oracle.h", "Makefile")]
    stamp = os.path.join(_HERE, "_build", "host_signature.txt")
    sig = _host_signature()
    built_for = open(stamp).read().strip() if os.path.exists(stamp) else ""
    stale = (not os.path.exists(_SO)) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src) or built_for != sig
    if force or stale:
        import shutil

        if shutil.which(os.environ.get("CC", "gcc")) is None and os.path.exists(_SO):
            return _SO  # no compiler on this box: keep the library that travelled (x86-64-v2 build is the fallback target)
        subprocess.run(["make", "-B", "-C", _HERE], check=True, stdout=subprocess.DEVNULL)
        with open(stamp, "w") as f:
            f.write(sig + "\n")
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, u8, u64 = C.c_void_p, C.c_int, C.c_uint8, C.c_uint64
        L.lo_world_default.argtypes = [vp]
        L.lo_world_empty.argtypes = [vp]
        L.lo_reconstruct.argtypes = [C.c_char_p, vp]
        L.lo_preserve.argtypes = [vp, C.c_char_p, C.c_size_t]
        L.lo_apply.argtypes = [vp, u8, u8, u8, u8, vp]
        L.lo_in_critical_endangerment.argtypes = [vp, i32]
        L.lo_is_being_leered_at_by.argtypes = [vp, u8, i32]
        for name in ("servant", "pony", "scholar", "cop", "princess", "figurehead", "service"):
            getattr(L, f"lo_{name}_lookahead").argtypes = [vp, i32, i32, vp, C.POINTER(i32)]
        L.lo_lookahead.argtypes = [vp, vp]
        L.lo_reckless_lookahead.argtypes = [vp, vp]
        L.lo_score.argtypes = [vp]
        L.lo_score.restype = C.c_float
        L.lo_score_batch.argtypes = [vp, C.c_size_t, vp]
        L.lo_perft.argtypes = [vp, i32]
        L.lo_perft.restype = u64
        L.lo_perft_divide.argtypes = [vp, i32, vp, i32, i32]
        L.lo_negamax.argtypes = [vp, i32]
        L.lo_negamax.restype = C.c_float
        L.lo_kickoff.argtypes = [vp, i32, i32, vp, vp, i32]
        L.lo_kickoff_quiet.argtypes = [vp, i32, i32, i32, vp, vp, i32]
        L.lo_negamax_quiet.argtypes = [vp, i32, i32]
        L.lo_negamax_quiet.restype = C.c_float
        L.lo_mix64.argtypes = [u64]
        L.lo_mix64.restype = u64
        L.lo_playout.argtypes = [u64, u64, vp, C.POINTER(i32)]
        L.lo_center_of_the_world.restype = u64
        L.lo_forward_contour.argtypes = [i32]
        L.lo_forward_contour.restype = u64
        L.lo_sideways_contour.argtypes = [i32]
        L.lo_sideways_contour.restype = u64
        _lib = L
    return _lib


class OraclePanic(RuntimeError):
    """The reference would have panicked (moral_panic! / unwrap on None) on this input."""


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _check(rc):
    if rc < 0:
        raise OraclePanic("reference panic path")
    return rc


def locale(rank, file):
    """space.rs:34 packing."""
    return (rank << 4) | file


def algebraic(name):
    """space.rs:53-63."""
    return locale(ord(name[1]) - 49, ord(name[0]) - 97)


def to_algebraic(loc):
    return "abcdefgh"[loc & 15] + str((loc >> 4) + 1)


def agent(team, job):
    return (team << 3) | job


def new_world():
    w = np.zeros((), dtype=WORLD_DTYPE)
    lib().lo_world_default(_ptr(w))
    return w


def empty_world():
    w = np.zeros((), dtype=WORLD_DTYPE)
    lib().lo_world_empty(_ptr(w))
    return w


def reconstruct(scan):
    w = np.zeros((), dtype=WORLD_DTYPE)
    _check(lib().lo_reconstruct(scan.encode(), _ptr(w)))
    return w


def preserve(w):
    w = np.ascontiguousarray(w)
    buf = C.create_string_buffer(128)
    _check(lib().lo_preserve(_ptr(w), buf, 128))
    return buf.value.decode()


def apply(w, team, job, whence, whither):
    w = np.ascontiguousarray(w)
    c = np.zeros((), dtype=COMMIT_DTYPE)
    _check(lib().lo_apply(_ptr(w), team, job, whence, whither, _ptr(c)))
    return c


def in_critical_endangerment(w, team):
    w = np.ascontiguousarray(w)
    return bool(_check(lib().lo_in_critical_endangerment(_ptr(w), team)))


def is_being_leered_at_by(w, loc, team):
    w = np.ascontiguousarray(w)
    return bool(_check(lib().lo_is_being_leered_at_by(_ptr(w), loc, team)))


def piece_lookahead(name, w, team, nihilistically):
    w = np.ascontiguousarray(w)
    out = np.zeros(MAX_COMMITS, dtype=COMMIT_DTYPE)
    n = C.c_int(0)
    _check(getattr(lib(), f"lo_{name}_lookahead")(_ptr(w), team, int(nihilistically), _ptr(out), C.byref(n)))
    return out[: n.value].copy()


def lookahead(w):
    w = np.ascontiguousarray(w)
    out = np.zeros(MAX_COMMITS, dtype=COMMIT_DTYPE)
    n = _check(lib().lo_lookahead(_ptr(w), _ptr(out)))
    return out[:n].copy()


def reckless_lookahead(w):
    w = np.ascontiguousarray(w)
    out = np.zeros(MAX_COMMITS, dtype=COMMIT_DTYPE)
    n = _check(lib().lo_reckless_lookahead(_ptr(w), _ptr(out)))
    return out[:n].copy()


def score(w):
    w = np.ascontiguousarray(w)
    return np.float32(lib().lo_score(_ptr(w)))


def score_batch(worlds):
    worlds = np.ascontiguousarray(worlds, dtype=WORLD_DTYPE)
    out = np.empty(len(worlds), dtype=np.float32)
    lib().lo_score_batch(_ptr(worlds), len(worlds), _ptr(out))
    return out


def perft(w, depth):
    w = np.ascontiguousarray(w)
    return int(lib().lo_perft(_ptr(w), depth))


def perft_divide(w, depth, threads=1):
    w = np.ascontiguousarray(w)
    per_root = np.zeros(MAX_COMMITS, dtype=np.uint64)
    n = _check(lib().lo_perft_divide(_ptr(w), depth, _ptr(per_root), MAX_COMMITS, threads))
    return per_root[:n].copy()


def negamax(w, depth, quiet=None):
    """mind.rs:271-364 TT-free; quiet = None or the quiescence extension (mind.rs:284-301)."""
    w = np.ascontiguousarray(w)
    return np.float32(lib().lo_negamax_quiet(_ptr(w), depth, -1 if quiet is None else quiet))


def kickoff(w, depth, nihilistically, threads=1, quiet=None):
    w = np.ascontiguousarray(w)
    roots = np.zeros(MAX_COMMITS, dtype=COMMIT_DTYPE)
    scores = np.zeros(MAX_COMMITS, dtype=np.float32)
    n = _check(lib().lo_kickoff_quiet(_ptr(w), depth, -1 if quiet is None else quiet, int(nihilistically), _ptr(roots), _ptr(scores), threads))
    return roots[:n].copy(), scores[:n].copy()


def playout(seed, index):
    w = np.zeros((), dtype=WORLD_DTYPE)
    plies = C.c_int(0)
    _check(lib().lo_playout(seed, index, _ptr(w), C.byref(plies)))
    return w, plies.value


def playout_batch(seed, first, n):
    out = np.zeros(n, dtype=WORLD_DTYPE)
    base = out.ctypes.data
    f = lib().lo_playout
    for i in range(n):
        _check(f(seed, first + i, base + 104 * i, None))
    return out
