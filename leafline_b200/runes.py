"""Preservation runes (the FEN-like position strings) on the host.

Mirrors `WorldState::reconstruct` (reference src/life.rs:428-507) and `WorldState::preserve`
(src/life.rs:293-360): four space-separated fields `board initiative eligibility passing-by`, extra
fields ignored, 'X' accepted as a field separator (life.rs:432).  Host-side glue only.
"""
import numpy as np

from ._abi import WORLD_DTYPE

_RUNES = "PNBRQKpnbrqk"  # identity.rs:119-147


def reconstruct(scan: str) -> np.ndarray:
    fields = scan.replace("X", " ").split(" ")
    if len(fields) < 4:
        raise ValueError("expected four rune fields (board, initiative, eligibility, passing-by)")
    world = np.zeros((), dtype=WORLD_DTYPE)
    planes = [0] * 12
    rank, file = 7, 0
    for rune in fields[0]:
        if rune == "/":
            file, rank = 0, rank - 1
        elif rune in "012345678":
            file += int(rune)
        else:
            agent = _RUNES.find(rune)
            if agent < 0:
                raise ValueError(f"tried to construct Agent from non-agent-preservation-rune {rune!r}")
            if not (0 <= rank < 8 and 0 <= file < 8):
                raise ValueError("board field runs off the world")
            planes[agent] |= 1 << (8 * rank + file)
            file += 1
    world["plane"] = np.array(planes, dtype=np.uint64)
    if fields[1][:1] == "w":
        world["initiative"] = 0
    elif fields[1][:1] == "b":
        world["initiative"] = 1
    else:
        raise ValueError("non-initiative-preserving rune")
    eligibility = 0
    for rune in fields[2]:
        if rune == "-":
            break
        bit = {"Q": 1, "K": 2, "q": 4, "k": 8}.get(rune)  # life.rs:178-181
        if bit is None:
            raise ValueError(f"non-eligibility rune {rune!r}")
        eligibility |= bit
    world["service_eligibility"] = eligibility
    if fields[3] == "-":
        world["passing_by"] = 0xFF
    else:
        f, r = ord(fields[3][0]) - 97, ord(fields[3][1]) - 49  # space.rs:53-63
        if not (0 <= r < 8 and 0 <= f < 8):
            raise ValueError("passing-by locale off the world")
        world["passing_by"] = (r << 4) | f
    return world


def preserve(world) -> str:
    planes = [int(x) for x in world["plane"]]
    rows = []
    for rank in range(7, -1, -1):
        row, run = "", 0
        for file in range(8):
            bit = 1 << (8 * rank + file)
            agent = next((a for a in range(12) if planes[a] & bit), None)
            if agent is None:
                run += 1
            else:
                if run:
                    row, run = row + str(run), 0
                row += _RUNES[agent]
        if run:
            row += str(run)
        rows.append(row)
    e = int(world["service_eligibility"])
    eligibility = "".join(r for r, bit in (("K", 2), ("Q", 1), ("k", 8), ("q", 4)) if e & bit) or "-"
    pb = int(world["passing_by"])
    passing = "-" if pb == 0xFF else "abcdefgh"[pb & 15] + str((pb >> 4) + 1)
    return "/".join(rows) + " " + "wb"[int(world["initiative"])] + " " + eligibility + " " + passing
