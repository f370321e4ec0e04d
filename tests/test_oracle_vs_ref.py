"""
Pins the oracle: the CPU restatement (oracle/sdft_oracle_impl.h) must be BIT-IDENTICAL to the
unmodified reference header compiled under oracle/_ref, for every type pair, window and latency, over
multi-call streams that cross the ring wrap and the modulation reset (c/src/sdft/sdft.h:562-598).
"""
import itertools

import numpy as np
import pytest

import oracle
from conftest import seed_of
from oracle import Oracle, Ref

pytestmark = pytest.mark.skipif(not oracle.have_ref(), reason="oracle/_ref not built (reference absent)")

TYPES = list(itertools.product(["f32", "f64"], repeat=2))


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("td,fd", TYPES)
@pytest.mark.parametrize("m", [1, 2, 3, 8, 37, 1000, 1024])
def test_tables_bit_identical(td, fd, m):
    for latency in (1.0, 0.5, 0.25, 0.9):
        o, r = Oracle(td, fd, m, 1, latency), Ref(td, fd, m, 1, latency)
        for a, b in zip(o.twiddles(), r.twiddles()):
            assert np.array_equal(_bits(a), _bits(b))


@pytest.mark.parametrize("td,fd", TYPES)
@pytest.mark.parametrize("window", [0, 1, 2, 3])
@pytest.mark.parametrize("latency", [1.0, 0.5])
def test_stream_bit_identical(td, fd, window, latency):
    rng = np.random.default_rng(seed_of(td, fd, window))
    for m in (1, 2, 3, 8, 37, 250):
        o, r = Oracle(td, fd, m, window, latency), Ref(td, fd, m, window, latency)
        for n in (1, 7, 100, 2 * m + 13, 5, 4 * m):
            x = rng.uniform(-1, 1, n)
            a, b = o.sdft(x), r.sdft(x)
            assert np.array_equal(_bits(a), _bits(b)), (m, n)
            assert np.array_equal(_bits(o.isdft(a)), _bits(r.isdft(b))), (m, n)
        so, sr = o.state(), r.state()
        assert so[0] == sr[0]
        for u, v in zip(so[1:], sr[1:]):
            assert np.array_equal(_bits(u), _bits(v))


def test_reset_matches():
    o, r = Oracle("f32", "f64", 16, 3, 0.5), Ref("f32", "f64", 16, 3, 0.5)
    x = np.linspace(-1, 1, 77)
    first = o.sdft(x)
    r.sdft(x)
    o.reset(); r.reset()
    a, b = o.sdft(x), r.sdft(x)
    assert np.array_equal(_bits(a), _bits(b))
    assert np.array_equal(_bits(a), _bits(first))


def test_closed_form_fft():
    """rows equal the first m bins of the length-2m FFT of the windowed last 2m
    samples, except the quirk bins next to m-1 (c/src/sdft/sdft.h:589-595)."""
    m, n = 64, 400
    rng = np.random.default_rng(3)
    x = rng.uniform(-1, 1, n)
    j = np.arange(2 * m)
    wins = {0: np.ones(2 * m), 1: 0.5 - 0.5 * np.cos(2 * np.pi * j / (2 * m)),
            2: 0.54 - 0.46 * np.cos(2 * np.pi * j / (2 * m)),
            3: 0.42 - 0.5 * np.cos(2 * np.pi * j / (2 * m)) + 0.08 * np.cos(4 * np.pi * j / (2 * m))}
    skip = {0: 0, 1: 1, 2: 1, 3: 2}
    xp = np.concatenate([np.zeros(2 * m), x])
    for w in range(4):
        d = Oracle("f64", "f64", m, w, 1.0).sdft(x)
        for t in (0, 5, 2 * m - 1, 2 * m, n - 1):
            seg = xp[t + 1:t + 1 + 2 * m]
            want = np.fft.fft(seg * wins[w])[:m] / (2 * m)
            good = m - skip[w]
            assert np.max(np.abs(d[t, :good] - want[:good])) < 1e-13
