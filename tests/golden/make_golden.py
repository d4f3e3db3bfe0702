#!/usr/bin/env python
"""Generates tests/golden/golden.json from the CPU oracle (oracle/leafline_oracle.c).

The Rust reference cannot be built or imported in this image, so these vectors are NOT outputs of the
reference itself: they are outputs of the restatement AFTER it passed every known-answer test the
reference holds for this path (tests/test_oracle_golden.py).  They freeze that state: a later change to
the oracle or to the CUDA path that moves any of them is caught by tests/test_golden_fixtures.py.

    python tests/golden/make_golden.py        # rewrites golden.json
"""
import hashlib
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))
import oracle as O  # noqa: E402
from positions import NAMED  # noqa: E402

PERFT_DEPTH = {"rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -": 4}


def commits_digest(commits):
    return hashlib.sha256(commits.tobytes()).hexdigest()


def main():
    out = {"seed": 0x1EAF11E, "positions": [], "playouts": {}}
    for fen in NAMED:
        w = O.reconstruct(fen)
        legal, reckless = O.lookahead(w), O.reckless_lookahead(w)
        depth = PERFT_DEPTH.get(fen, 3)
        out["positions"].append({
            "fen": fen,
            "preserved": O.preserve(w),
            "score_bits": int(np.float32(O.score(w)).view(np.uint32)),
            "legal": len(legal), "reckless": len(reckless),
            "legal_sha256": commits_digest(legal), "reckless_sha256": commits_digest(reckless),
            "legal_moves": [f"{O.to_algebraic(c['whence'])}{O.to_algebraic(c['whither'])}" + ("" if c["ascension"] == 0xFF else "=" + "PNBRQK"[c["ascension"] & 7]) for c in legal],
            "perft": [O.perft(w, d) for d in range(1, depth + 1)],
            "negamax_bits": [int(np.float32(O.negamax(w, d)).view(np.uint32)) for d in range(0, 3)],
        })
    n = 256
    worlds = O.playout_batch(out["seed"], 0, n)
    out["playouts"] = {
        "n": n,
        "worlds_sha256": hashlib.sha256(worlds.tobytes()).hexdigest(),
        "fens": [O.preserve(w) for w in worlds[:16]],
        "score_bits": [int(x) for x in O.score_batch(worlds).view(np.uint32)],
        "legal_counts": [len(O.lookahead(w)) for w in worlds],
        "reckless_counts": [len(O.reckless_lookahead(w)) for w in worlds],
    }
    with open(os.path.join(HERE, "golden.json"), "w") as f:
        json.dump(out, f, indent=0, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "golden.json"), os.path.getsize(os.path.join(HERE, "golden.json")), "bytes")


if __name__ == "__main__":
    main()
