"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol the header
declares, and fails loudly (no CPU fallback) without a GPU.  No compute calls here."""
import ctypes as C
import os
import re

import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    entry.build()
    from leafline_b200 import _abi

    return _abi.load()


def header_symbols():
    text = open(os.path.join(ROOT, "include", "leafline_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ll_[a-z_0-9]+)\s*\(", text)))


def test_header_and_binding_agree(lib):
    from leafline_b200 import _abi

    declared = header_symbols()
    assert declared, "no symbols parsed from the header"
    assert sorted(_abi.SIGNATURES) == declared
    for name in declared:
        assert hasattr(lib, name), f"{name} is declared in include/leafline_b200.h but not exported"


def c_parameter_counts():
    text = open(os.path.join(ROOT, "include", "leafline_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return {name: (0 if args.strip() in ("", "void") else args.count(",") + 1)
            for name, args in re.findall(r"\b(ll_[a-z_0-9]+)\s*\(([^)]*)\)\s*;", text)}


def test_rust_shim_declares_the_header(lib):
    # host/rust/b200.rs is shipped as source (no Rust toolchain here): every extern it declares must be an entry point
    # of the header with the same number of parameters, and the repr(C) structs must have the header's sizes
    text = open(os.path.join(ROOT, "host", "rust", "b200.rs")).read()
    block = text[text.index('extern "C" {'):]
    block = block[:block.index("\n}\n")]
    declared = re.findall(r"fn (ll_[a-z_0-9]+)\(([^)]*)\)", block)
    counts = c_parameter_counts()
    assert len(declared) >= 12
    for name, args in declared:
        assert name in counts, f"{name} is not in include/leafline_b200.h"
        n = 0 if not args.strip() else args.count(",") + 1
        assert n == counts[name], (name, n, counts[name])
        assert hasattr(lib, name)
    assert "plane: [u64; 12]" in text and "pad: [u8; 5]" in text and "variation: [LlPatch; LL_MAX_VARIATION]" in text
    assert "const LL_MAX_VARIATION: usize = 16;" in text


def test_struct_sizes_match_header():
    from leafline_b200 import _abi

    assert _abi.WORLD_DTYPE.itemsize == 104
    assert _abi.COMMIT_DTYPE.itemsize == 112
    assert _abi.COMMIT_DTYPE.fields["tree"][1] == 8
    assert C.sizeof(_abi.Soa) == 32


def test_no_cpu_fallback_without_a_device(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from leafline_b200 import _abi

    ctx = C.c_void_p()
    rc = lib.ll_init(0, C.byref(ctx))
    assert rc == _abi.ERR_NO_DEVICE
    assert b"no CPU fallback" in lib.ll_last_error()
    import leafline_b200 as ll

    with pytest.raises(ll.LeaflineError):
        ll.Engine(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "leafline_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp", ".hpp")):
                text = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in text and "leafline_oracle" not in text, f


def test_runes_round_trip_matches_oracle():
    import oracle as O
    import leafline_b200 as ll

    fens = [
        "rnbqkbnr/pppppppp/8/8/8/8/PPPPPPPP/RNBQKBNR w KQkq -",
        "rnbqkbnr/pppppppp/8/8/4P3/8/PPPP1PPP/RNBQKBNR b KQkq e3",
        "rnbqkbnr/pp1ppppp/8/2p5/4P3/8/PPPP1PPP/RNBQKBNR w KQkq c6",
        "r3k2r/p1ppqpb1/bn2pnp1/3PN3/1p2P3/2N2Q1p/PPPBBPPP/R3K2R w KQkq - 0 1",
        "8/q1P1k/8/8/8/8/6PP/7K w - -",
        "8/8/4k3/8/b7/8/8/R3KN1R w Q - 0 1",
    ]
    for fen in fens:
        a, b = ll.reconstruct(fen), O.reconstruct(fen)
        assert a.tobytes() == b.tobytes()
        assert ll.preserve(a) == O.preserve(b)
    assert ll.new_world().tobytes() == O.new_world().tobytes()
