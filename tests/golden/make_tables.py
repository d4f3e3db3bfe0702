#!/usr/bin/env python
"""Generates tests/golden/reference_tables.json by RUNNING the reference's own table generator.

`furniture/tablemaker.py` is the one piece of the reference that executes in this image (pure Python 3): it is
imported from /root/reference as it lies there, unmodified, and every value it would write into `motion.rs` /
`landmark.rs` is recorded -- the two 64-entry movement tables (tablemaker.py:41-62) and the landmark masks
(tablemaker.py:23-38, 91-109).  The fixture travels to the GPU box (the reference tree does not);
tests/test_oracle_golden.py regenerates it whenever /root/reference is present and compares it with the oracle's
LO_* tables, tests/test_gpu_parity.py compares it with the tables resident on the device.

    python tests/golden/make_tables.py        # rewrites reference_tables.json
"""
import importlib.util
import json
import os

HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE = os.environ.get("LEAFLINE_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "reference_tables.json")


def reference_tables(reference=REFERENCE):
    """the values tablemaker.main() writes, taken from the module's own functions"""
    path = os.path.join(reference, "furniture", "tablemaker.py")
    spec = importlib.util.spec_from_file_location("leafline_reference_tablemaker", path)
    tm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(tm)  # defines functions only; main() is not called (it writes into src/)
    return {
        "source": "furniture/tablemaker.py (zackmdavis/Leafline), imported unmodified",
        "pony_movement_table": [int(x) for x in tm.universal_distribution(tm.PONY_OPTIONS)],
        "figurehead_movement_table": [int(x) for x in tm.universal_distribution(tm.FIGUREHEAD_OPTIONS)],
        "center_of_the_world": int(tm.center_of_the_world()),
        "low_seventh_heaven": int(tm.forward_contour(1)),
        "low_colonelcy": int(tm.forward_contour(2)),
        "high_seventh_heaven": int(tm.forward_contour(6)),
        "high_colonelcy": int(tm.forward_contour(5)),
        "files": [int(tm.sideways_contour(f)) for f in range(8)],
    }


if __name__ == "__main__":
    with open(OUT, "w") as f:
        json.dump(reference_tables(), f, indent=1)
        f.write("\n")
    print("wrote", OUT)
