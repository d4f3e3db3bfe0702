"""The compiled host side: host/leafline.hpp mirrors the reference's Rust interface over the C ABI and
host/selfcheck.cpp asks the reference's own known answers of it (a C++ program, no Python in the loop)."""
import os
import subprocess

import pytest

import __graft_entry__ as entry

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BINARY = os.path.join(ROOT, "host", "selfcheck")


@pytest.fixture(scope="module")
def binary():
    entry.build()  # builds the library, then `make -C host`
    assert os.path.exists(BINARY)
    return BINARY


def test_host_program_builds_and_refuses_to_run_without_a_gpu(binary):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    run = subprocess.run([binary], capture_output=True, text=True)
    assert run.returncode != 0
    assert "no CPU fallback" in run.stderr  # the rune round trips pass first; constructing the Mind is what fails


@pytest.mark.gpu
def test_host_program_reproduces_the_reference_known_answers(binary):
    run = subprocess.run([binary], capture_output=True, text=True, timeout=300)
    assert run.returncode == 0, run.stderr
    assert "host selfcheck ok" in run.stdout
