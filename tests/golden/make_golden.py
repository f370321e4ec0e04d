"""
Regenerates the committed golden vectors.  Runs ONLY in the authoring container, where the reference
tree is mounted at /root/reference; the GPU box never executes this script.

  python tests/golden/make_golden.py

Sources of truth:
  * the reference's C header, compiled unmodified by oracle/Makefile (oracle.Ref)       -> c_ref_*.npz
  * the reference's Python class python/src/sdft/sdft.py imported from /root/reference  -> py_ref.npz,
                                                                                           py_convolve.npz
  * the reference's own integration-test input test/test.wav and test parameters
    (test/main.sh:3-6: DFTSIZE=1000 HOPSIZE=100 hann latency 1; BASELINE.json config 1: m=1024)
                                                                                          -> testwav.npz
"""
import hashlib
import os
import sys
import wave

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import Ref, build  # noqa: E402

REF = "/root/reference"


def read_pcm24(path):
    """24-bit PCM mono -> int32 sample values (chunk-aware via the wave module)."""
    with wave.open(path, "rb") as f:
        assert f.getsampwidth() == 3 and f.getnchannels() == 1
        sr = f.getframerate()
        raw = np.frombuffer(f.readframes(f.getnframes()), dtype=np.uint8).reshape(-1, 3).astype(np.int32)
    v = raw[:, 0] | (raw[:, 1] << 8) | (raw[:, 2] << 16)
    v = np.where(v & 0x800000, v - (1 << 24), v).astype(np.int32)
    return v, sr


def small_cases():
    rng = np.random.default_rng(0x5DF7)
    out = {}
    idx = 0
    for td in ("f32", "f64"):
        for fd in ("f32", "f64"):
            for m in (3, 8, 19):
                for window in range(4):
                    for latency in (1.0, 0.5):
                        calls = [1, 7, 2 * m + 13, 5, 12]
                        r = Ref(td, fd, m, window, latency)
                        xs, ds, ys = [], [], []
                        for n in calls:
                            x = rng.uniform(-1, 1, n).astype(r.td_np)
                            d = r.sdft(x)
                            xs.append(x); ds.append(d); ys.append(r.isdft(d))
                        key = "case%03d" % idx
                        out[key + "_meta"] = np.array([m, window, int(latency * 1000), len(calls)] + calls, np.int64)
                        out[key + "_types"] = np.array([td, fd])
                        out[key + "_x"] = np.concatenate(xs)
                        out[key + "_dft"] = np.concatenate(ds)
                        out[key + "_y"] = np.concatenate(ys)
                        a, s = r.twiddles()
                        out[key + "_tw"] = a
                        out[key + "_tws"] = s
                        idx += 1
    out["count"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "c_ref_small.npz"), **out)
    print("c_ref_small.npz:", idx, "cases")


def table_cases():
    """Twiddle tables at the BASELINE sizes (float tables are parity-critical)."""
    out = {}
    for fd in ("f32", "f64"):
        for m in (512, 1000, 1024, 2048, 4096):
            for latency in (1.0, 0.5):
                r = Ref("f32", fd, m, 1, latency)
                a, s = r.twiddles()
                out["tw_%s_%d_%d" % (fd, m, int(latency * 1000))] = a
                out["tws_%s_%d_%d" % (fd, m, int(latency * 1000))] = s
    np.savez_compressed(os.path.join(HERE, "c_ref_tables.npz"), **out)
    print("c_ref_tables.npz:", len(out), "tables")


def python_cases():
    sys.path.insert(0, os.path.join(REF, "python", "src"))
    from sdft import SDFT  # the reference's NumPy class
    rng = np.random.default_rng(0x5DF8)
    out = {}
    idx = 0
    for m in (16, 23):
        for window in ("boxcar", "hann", "hamming", "blackman"):
            for latency in (1, 0.5):
                s = SDFT(m, window, latency)
                calls = [11, 2 * m + 3, 9]
                xs, ds, ys = [], [], []
                for n in calls:
                    x = rng.uniform(-1, 1, n)
                    d = s.sdft(x)
                    xs.append(x); ds.append(d); ys.append(s.isdft(d))
                key = "case%03d" % idx
                out[key + "_meta"] = np.array([m, ["boxcar", "hann", "hamming", "blackman"].index(window),
                                               int(latency * 1000), len(calls)] + calls, np.int64)
                out[key + "_x"] = np.concatenate(xs)
                out[key + "_dft"] = np.concatenate(ds)
                out[key + "_y"] = np.concatenate(ys)
                idx += 1
    out["count"] = np.array(idx)
    np.savez_compressed(os.path.join(HERE, "py_ref.npz"), **out)
    print("py_ref.npz:", idx, "cases")


def python_convolve_cases():
    """SDFT.convolve of the reference's Python class on random un-windowed matrices -> py_convolve.npz"""
    sys.path.insert(0, os.path.join(REF, "python", "src"))
    from sdft import SDFT
    rng = np.random.default_rng(0x5DF9)
    out = {}
    for m in (8, 37):
        x = rng.uniform(-1, 1, (6, m)) + 1j * rng.uniform(-1, 1, (6, m))
        out["x_m%d" % m] = x
        for window in ("boxcar", "hann", "hamming", "blackman"):
            out["y_m%d_%s" % (m, window)] = SDFT(m, window, 1).convolve(x)
    np.savez_compressed(os.path.join(HERE, "py_convolve.npz"), **out)
    print("py_convolve.npz written")


def testwav_cases():
    pcm, sr = read_pcm24(os.path.join(REF, "test", "test.wav"))
    sha = hashlib.sha256(open(os.path.join(REF, "test", "test.wav"), "rb").read()).hexdigest()
    x = (pcm.astype(np.float64) / 8388608.0).astype(np.float32)  # dr_wav s24 -> f32 scaling
    out = {"pcm24": pcm, "sr": np.array(sr), "sha256": np.array(sha)}

    # (1) the reference's own integration test parameters: m=1000, hop=100, hann, latency 1; the test
    #     keeps the DFT row of the first sample of every hop (test/test.c:79-82) and the resynthesis.
    m, hop = 1000, 100
    nh = 24
    r = Ref("f32", "f64", m, 1, 1.0)
    rows, ys = [], []
    for h in range(nh):
        d = r.sdft(x[h * hop:(h + 1) * hop])
        rows.append(d[0].copy()); ys.append(r.isdft(d))
    out["t1000_rows"] = np.stack(rows)
    out["t1000_y"] = np.concatenate(ys)

    # (2) BASELINE config 1: m=1024, hann, f32 TD / f64 FD, latency 1, whole signal in 4096-sample calls.
    #     Keep the last row of every 8th call, every 8th synthesized sample and the sha256 of all of them.
    m, call = 1024, 4096
    n = (x.size // call) * call
    r = Ref("f32", "f64", m, 1, 1.0)
    rows, ys, row_t = [], [], []
    for c in range(n // call):
        d = r.sdft(x[c * call:(c + 1) * call])
        ys.append(r.isdft(d))
        if c % 8 == 0:
            rows.append(d[call - 1].copy()); row_t.append(c * call + call - 1)
    y = np.concatenate(ys)
    out["c1_n"] = np.array(n)
    out["c1_row_t"] = np.array(row_t, np.int64)
    out["c1_rows"] = np.stack(rows)
    out["c1_y_stride"] = np.array(8)
    out["c1_y_strided"] = y[::8].copy()
    out["c1_y_sha256"] = np.array(hashlib.sha256(y.tobytes()).hexdigest())
    delay = m - 1
    e = y[delay:].astype(np.float64) - x[:n - delay].astype(np.float64)
    out["c1_snr_db"] = np.array(10 * np.log10(np.mean(x[:n - delay].astype(np.float64) ** 2) / np.mean(e ** 2)))
    np.savez_compressed(os.path.join(HERE, "testwav.npz"), **out)
    print("testwav.npz: n =", n, "snr =", float(out["c1_snr_db"]), "dB")


def chirp_late_rows():
    """BASELINE config 3 at FULL length: the 2^26-sample chirp, m = 2048, float FD, hann, latency 0.5, run
    continuously through the compiled reference (the parity build, ~15 CPU-minutes on one core; rows are produced
    in 4096-sample tiles and dropped).  Kept: 4 rows right after the 8-way time-shard boundaries 2, 4, 6 and 7
    and the last 4 rows of the signal -- where the float accumulators have walked longest -- plus the samples
    the reference synthesizes from them.  Pins the long-run float path: tests/test_gpu_configs.py compares the
    continuous GPU run and the halo-primed shards with these rows at 1e-4."""
    sys.path.insert(0, os.path.join(HERE, "..", ".."))
    from sdft_b200 import workloads
    n, m, tile, keep = 1 << 26, 2048, 4096, 4
    probes = [k * (n // 8) for k in (2, 4, 6, 7)] + [n - keep]
    r = Ref("f32", "f32", m, 1, 0.5)
    rows, ys = {}, {}
    pos = 0
    import time
    t0 = time.time()
    while pos < n:
        x = workloads.chirp(n, pos, tile)
        d = r.sdft(x)
        for p in probes:
            if pos <= p < pos + tile:
                assert p + keep <= pos + tile
                rows[p] = d[p - pos:p - pos + keep].copy()
                ys[p] = r.isdft(rows[p])
        pos += tile
        if (pos // tile) % 1024 == 0:
            print("chirp: %d / %d samples, %.0f s" % (pos, n, time.time() - t0), flush=True)
    out = {"n": np.array(n), "m": np.array(m), "probes": np.array(probes, np.int64),
           "rows": np.stack([rows[p] for p in probes]), "y": np.stack([ys[p] for p in probes]),
           "final_cursor": np.array(r.state()[0])}
    np.savez_compressed(os.path.join(HERE, "c3_chirp_late.npz"), **out)
    print("c3_chirp_late.npz written in %.0f s" % (time.time() - t0))


if __name__ == "__main__":
    build(want_ref=True)
    if "--chirp" in sys.argv:          # the long one, on request only
        chirp_late_rows()
        sys.exit(0)
    small_cases()
    table_cases()
    python_cases()
    python_convolve_cases()
    testwav_cases()
