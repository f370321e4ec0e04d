"""
CPU checks of the test/bench infrastructure itself: the synthetic inputs of the BASELINE configurations,
the closed form used for drift checks, and the oracle helpers that let the full-size GPU tests walk long
signals (advance / clone / roundtrip must not change a single bit of the oracle's state or rows).
"""
import numpy as np
import pytest

from oracle import Oracle
from sdft_b200 import workloads

WINDOWS = ("boxcar", "hann", "hamming", "blackman")


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


def test_stream_is_split_invariant_and_bounded():
    a = workloads.stream(0, 10000)
    b = np.concatenate([workloads.stream(0, 777), workloads.stream(777, 10000 - 777)])
    assert np.array_equal(a, b)
    assert a.dtype == np.float32 and np.abs(a).max() <= 0.75
    far = workloads.stream((1 << 30) - 4096, 4096, dtype=np.float64)
    assert np.isfinite(far).all() and np.abs(far).max() <= 0.75


def test_chirp_and_channels_are_reproducible():
    x = workloads.chirp(1 << 16)
    assert x.dtype == np.float32 and x[0] == 0.0
    assert np.array_equal(x[1000:2000], workloads.chirp(1 << 16, 1000, 1000))
    assert not np.array_equal(workloads.channel_noise(0, 64), workloads.channel_noise(1, 64))
    assert np.array_equal(workloads.channel_noise(5, 64), workloads.channel_noise(5, 64))


@pytest.mark.parametrize("window", WINDOWS)
def test_closed_form_matches_oracle_away_from_mirror_bins(window):
    """rows are the first m bins of the 2m-point FFT of the windowed last 2m samples,
    except the last bin (hann, hamming) / last two bins (blackman), which the reference distorts by
    mirroring its upper halo about bin m-1 (c/src/sdft/sdft.h:589-595)."""
    m, n = 64, 1000
    x = workloads.white_noise(n).astype(np.float64)
    rows = Oracle("f64", "f64", m, window, 1.0).sdft(x)
    for t in (2 * m - 1, 500, n - 1):
        want = workloads.closed_form_row(x[t + 1 - 2 * m:t + 1], window)
        good = m if window == "boxcar" else (m - 2 if window == "blackman" else m - 1)
        assert np.abs(rows[t] - want)[:good].max() <= 1e-13
        if window != "boxcar":
            assert np.abs(rows[t] - want)[good:].max() > 1e-6


@pytest.mark.parametrize("td,fd", [("f32", "f64"), ("f32", "f32"), ("f64", "f64")])
def test_oracle_walk_helpers_are_bit_exact(td, fd):
    m = 37
    x = workloads.white_noise(5 * m + 11, seed=9)
    full = {w: Oracle(td, fd, m, w, 0.5).sdft(x) for w in WINDOWS}
    walk = Oracle(td, fd, m, "boxcar", 0.5)
    cut = 2 * m + 5
    walk.advance(x[:cut])
    for w in WINDOWS:
        clone = walk.clone(w)
        assert np.array_equal(_bits(clone.sdft(x[cut:])), _bits(full[w][cut:]))
    # the walk itself is untouched by its clones
    ref = Oracle(td, fd, m, "boxcar", 0.5)
    ref.sdft(x[:cut])
    for a, b in zip(walk.state(), ref.state()):
        assert np.array_equal(_bits(np.asarray(a)), _bits(np.asarray(b)))
    hann = Oracle(td, fd, m, "hann", 0.5)
    y = Oracle(td, fd, m, "hann", 0.5).roundtrip(x)
    assert np.array_equal(_bits(y), _bits(hann.isdft(hann.sdft(x))))


def test_snr_definition():
    x = np.sin(np.arange(5000) * 0.01)
    y = np.concatenate([np.zeros(7), x])[:5000]
    assert workloads.snr_db(x, y, 7) > 200
    assert abs(workloads.snr_db(x, y + 0.1 * np.concatenate([np.zeros(7), x])[:5000], 7) - 20.0) < 1e-6
