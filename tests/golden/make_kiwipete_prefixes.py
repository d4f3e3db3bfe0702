#!/usr/bin/env python
"""Generates tests/golden/kiwipete_prefix_perft5.json: the CPU oracle's perft(5) below a seeded sample of 2-ply prefixes
of Kiwipete (BASELINE.json configs[4]: perft(7) of a tactics-heavy position).

The faithful O(moves^2) oracle cannot count the whole depth-7 tree (3.7e11 leaves); SURVEY.md 8(d) prescribes a
hierarchical check instead: the oracle recounts a seeded sample of 2-ply prefixes at sub-depth 5 (about 1.8e8 leaves
each, a minute apiece on eight host threads), the GPU must reproduce every one of them, and the whole-tree count is
then held by GPU self-consistency (sum of the 48 root children, one GPU vs many, launch chain vs level by level).
bench.py's config5 block and tests/test_gpu_parity.py compare against this fixture.

    python tests/golden/make_kiwipete_prefixes.py [n_samples]      # rewrites the fixture (about 20 minutes)
"""
import json
import os
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import oracle as O  # noqa: E402

KIWIPETE = "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq -"
SEED = 0x1EAF11E
OUT = os.path.join(HERE, "kiwipete_prefix_perft5.json")


def main(n_samples=16):
    O.build()
    root = O.reconstruct(KIWIPETE)
    prefixes = []
    for i, c1 in enumerate(O.lookahead(root)):
        for j, c2 in enumerate(O.lookahead(c1["tree"])):
            prefixes.append((i, j, c2["tree"].copy()))
    assert len(prefixes) == 2038  # the reference's perft(2) of Kiwipete
    picks, h = [], SEED
    while len(picks) < n_samples:
        h = int(O.lib().lo_mix64(h))
        k = h % len(prefixes)
        if k not in picks:
            picks.append(k)
    samples = []
    threads = os.cpu_count() or 1
    for k in picks:
        i, j, w = prefixes[k]
        t0 = time.time()
        leaves = int(O.perft_divide(w, 5, threads=threads).sum())
        samples.append({"unit": k, "root_move": i, "reply": j, "runes": O.preserve(w), "perft5": leaves})
        print(k, O.preserve(w), leaves, f"{time.time() - t0:.0f}s", flush=True)
    with open(OUT, "w") as f:
        json.dump({"position": KIWIPETE, "seed": SEED, "sub_depth": 5, "n_prefixes": len(prefixes),
                   "source": "oracle/leafline_oracle.c (lo_perft_divide), tests/golden/make_kiwipete_prefixes.py", "samples": samples}, f, indent=1)
        f.write("\n")


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 16)
