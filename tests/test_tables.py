"""Host logic on the CPU: the twiddle tables the library uploads (sdft_b200/csrc/sdft_tables.hpp, compiled with g++
on its own) must be BIT-identical to the reference's (through the oracle, which is pinned to the compiled reference
header): float-FD parity depends on it."""
import os
import subprocess

import numpy as np
import pytest

from oracle import Oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("tables") / "tables_dump")
    subprocess.run(["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-Werror",
                    "-I", os.path.join(ROOT, "sdft_b200", "csrc"),
                    os.path.join(ROOT, "tests", "drivers", "tables_dump.cpp"), "-o", exe], check=True)
    return exe


@pytest.mark.parametrize("fd", ["f32", "f64"])
@pytest.mark.parametrize("m", [1, 2, 3, 8, 37, 1000, 1024, 4096])
def test_tables_bit_identical_to_the_reference(dumper, fd, m):
    cdt = np.complex64 if fd == "f32" else np.complex128
    for latency in (1.0, 0.5, 0.3):
        raw = subprocess.run([dumper, fd, str(m), repr(latency)], capture_output=True, check=True).stdout
        got = np.frombuffer(raw, cdt).reshape(2, m)
        want = Oracle("f32", fd, m, "hann", latency).twiddles()
        for g, w in zip(got, want):
            assert np.array_equal(g.view(np.uint8), np.ascontiguousarray(w).view(np.uint8)), (fd, m, latency)
