"""Host logic on the CPU: the synthesis weights with the window folded in (sdft_b200/csrc/sdft_weights.hpp, compiled
with g++ on its own).  For random spectra the folded form  sum_b (A_b Re aux_b + B_b Im aux_b)  must equal what the
reference computes in two steps: mirror the halo cells (c/src/sdft/sdft.h:589-595, in its assignment order), apply
the window taps (sdft.h:350-402), then weigh the bins as sdft_isdft does (sdft.h:639-652)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TAPS = {0: [0, 0, 1.0, 0, 0], 1: [0, -0.25, 0.5, -0.25, 0], 2: [0, -0.23, 0.54, -0.23, 0], 3: [0.04, -0.25, 0.42, -0.25, 0.04]}


@pytest.fixture(scope="module")
def dumper(tmp_path_factory):
    exe = str(tmp_path_factory.mktemp("weights") / "weights_dump")
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I", os.path.join(ROOT, "sdft_b200", "csrc"),
                    os.path.join(ROOT, "tests", "drivers", "weights_dump.cpp"), "-o", exe], check=True)
    return exe


def reference_two_step(aux, v, window):
    """windowed rows exactly as the reference forms them, then the weighted real sum of sdft_isdft"""
    m = aux.size
    ext = np.zeros(m + 4, np.complex128)
    ext[2:m + 2] = aux
    for i in (1, 2):                                    # sdft.h:589-595, same interleaved order
        ext[2 - i] = np.conj(ext[2 + i])
        ext[(m + 1) + i] = np.conj(ext[(m + 1) - i])
    w = 1.0 / (2 * m)
    rows = np.array([sum(TAPS[window][j + 2] * ext[k + 2 + j] for j in range(-2, 3)) * w for k in range(m)])
    return float(np.sum((rows * v).real))


@pytest.mark.parametrize("window", [0, 1, 2, 3])
@pytest.mark.parametrize("m", [1, 2, 3, 4, 8, 37])
def test_folded_weights_equal_window_then_weigh(dumper, m, window):
    rng = np.random.default_rng(100 * m + window)
    for unit in (True, False):
        v = np.where(np.arange(m) % 2, -1.0, 1.0).astype(np.complex128) if unit else \
            rng.uniform(-1, 1, m) + 1j * rng.uniform(-1, 1, m)
        prescale = 1.0 if unit else 0.37
        blob = struct.pack("qqd", m, window, prescale) + v.real.astype(np.float64).tobytes() + v.imag.astype(np.float64).tobytes()
        raw = subprocess.run([dumper], input=blob, capture_output=True, check=True).stdout
        ab = np.frombuffer(raw[:16 * m], np.float64).reshape(m, 2) * prescale
        assert raw[16 * m] == (1 if unit else 0) or m == 1
        for _ in range(5):
            aux = rng.uniform(-1, 1, m) + 1j * rng.uniform(-1, 1, m)
            folded = float(np.sum(ab[:, 0] * aux.real + ab[:, 1] * aux.imag))
            assert abs(folded - reference_two_step(aux, v, window)) <= 1e-13, (m, window, unit)
