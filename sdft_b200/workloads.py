"""
Synthetic inputs of the BASELINE.json configurations, shared by bench.py and the parity
tests so that both sides of every comparison see the same samples.  Pure NumPy, nothing here computes
the transform.

  config 2   2^20 samples of white noise in [-1, 1), float32
  config 3   2^26-sample linear chirp 0 -> 0.25 fs, phase evaluated in float64, cast to float32
  config 4   512 independent white-noise channels, one generator per channel
  config 5   endless stream: 0.5 sin(2 pi 0.01 t) + 0.25 u(t), u a counter-based hash in [-1, 1)
"""
import numpy as np

SEED_C2 = 0x5DF70002
SEED_C4 = 0x5DF70004


def white_noise(n, seed=SEED_C2, dtype=np.float32):
    return np.random.default_rng(seed).uniform(-1, 1, n).astype(dtype)


def chirp(n, begin=0, count=None, dtype=np.float32):
    """x[t] = sin(pi * 0.25 * t^2 / n) for t in [begin, begin + count): instantaneous frequency rises
    linearly from 0 to 0.25 fs over the n samples of the whole signal."""
    count = n - begin if count is None else count
    t = np.arange(begin, begin + count, dtype=np.float64)
    return np.sin(np.pi * 0.25 * t * t / float(n)).astype(dtype)


def channel_noise(channel, n, dtype=np.float32):
    return np.random.default_rng([SEED_C4, int(channel)]).uniform(-1, 1, n).astype(dtype)


def _splitmix64(z):
    z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
    z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    return z ^ (z >> np.uint64(31))


def stream(begin, count, dtype=np.float32):
    """Samples [begin, begin + count) of the endless stream; any call split gives the same samples."""
    t = np.arange(begin, begin + count, dtype=np.uint64)
    with np.errstate(over="ignore"):
        h = _splitmix64(t)
    u = (h >> np.uint64(11)).astype(np.float64) * (2.0 / (1 << 53)) - 1.0
    tone = 0.5 * np.sin(2 * np.pi * 0.01 * (t % np.uint64(100)).astype(np.float64))   # period 100 samples, exact
    return (tone + 0.25 * u).astype(dtype)


def window_taps(window, n):
    """The periodic time-domain window of length n that the reference's frequency-domain taps
    implement (c/src/sdft/sdft.h:350-402): boxcar, hann, hamming, blackman."""
    j = np.arange(n, dtype=np.float64)
    c = np.cos(2 * np.pi * j / n)
    if window in (1, "hann"):
        return 0.5 - 0.5 * c
    if window in (2, "hamming"):
        return 0.54 - 0.46 * c
    if window in (3, "blackman"):
        return 0.42 - 0.5 * c + 0.08 * np.cos(4 * np.pi * j / n)
    return np.ones(n)


def closed_form_row(last_2m_samples, window):
    """dft[t][k] = (1/2m) sum_j x[t-2m+1+j] w[j] e^{-2 pi i j k / 2m}, k < m.  Valid
    away from the last bin (hann/hamming) or the last two bins (blackman), where the reference mirrors
    its upper halo about bin m-1 (c/src/sdft/sdft.h:589-595)."""
    seg = np.asarray(last_2m_samples, dtype=np.float64)
    n = seg.size
    return np.fft.fft(seg * window_taps(window, n))[: n // 2] / n


def snr_db(x, y, delay):
    """Reconstruction SNR as python/examples/latency.py:30-56 defines it: the synthesized signal lags
    the input by `delay` = int((m - 1) * latency) samples."""
    x = np.asarray(x, dtype=np.float64)
    y = np.asarray(y, dtype=np.float64)
    xd = x[: x.size - delay]
    e = y[delay:] - xd
    return 10 * np.log10(np.mean(xd ** 2) / np.mean(e ** 2))
