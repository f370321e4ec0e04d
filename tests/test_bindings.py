"""
The C shim (include/c/sdft/sdft.h) and the C++ template (include/cpp/sdft/sdft.h) over the C-ABI
library: they must compile and link for every type combination on the CPU box, and on the GPU the
reference-shaped drivers (test/test.c, test/test.cpp pattern: hop-wise sdft_n + isdft_n) must agree
with the oracle the way the reference's own integration test demands (test/main.py:70,78: allclose).
"""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DRV = os.path.join(ROOT, "tests", "drivers")
OUT = os.path.join(ROOT, "tests", "drivers", "_build")
LIBDIR = os.path.join(ROOT, "sdft_b200")


def _compile(src, exe, lang, defines=()):
    from sdft_b200 import build
    build.build()
    os.makedirs(OUT, exist_ok=True)
    inc = os.path.join(ROOT, "include", "c" if lang == "c" else "cpp")
    cc = ["gcc", "-std=gnu99"] if lang == "c" else ["g++", "-std=c++11"]
    cmd = cc + ["-O2", "-Wall", "-Werror", "-I", inc] + ["-D" + d for d in defines] + [
        src, "-o", exe, "-L", LIBDIR, "-lsdft_b200", "-Wl,-rpath," + LIBDIR]
    subprocess.run(cmd, check=True)
    return exe


@pytest.mark.parametrize("td", ["SDFT_TD_FLOAT", "SDFT_TD_DOUBLE", None])
@pytest.mark.parametrize("fd", ["SDFT_FD_FLOAT", "SDFT_FD_DOUBLE", None])
@pytest.mark.parametrize("nocomplex", [False, True])
def test_c_shim_compiles_for_every_type_selection(td, fd, nocomplex, tmp_path):
    src = tmp_path / "t.c"
    src.write_text("""
#include <sdft/sdft.h>
int main(void)
{
  sdft_t* s = sdft_alloc_custom(8, sdft_window_blackman, 0.5);
  sdft_td_t x[4] = {0}; sdft_fdx_t d[32]; sdft_fdx_t* rows[4] = {d, d + 8, d + 16, d + 24};
  const sdft_fdx_t* crows[4] = {d, d + 8, d + 16, d + 24};
  if (!s) return 0;
  sdft_sdft(s, x[0], d); sdft_sdft_n(s, 4, x, d); sdft_sdft_nd(s, 4, x, rows);
  x[0] = sdft_isdft(s, d); sdft_isdft_n(s, 4, d, x); sdft_isdft_nd(s, 4, crows, x);
  sdft_reset(s);
  return (int)(sdft_size(s) + (size_t)sdft_window(s) + (size_t)sdft_latency(s)) * 0 + (sdft_free(s), 0);
}
""")
    defs = [d for d in (td, fd) if d] + (["SDFT_NO_COMPLEX_H"] if nocomplex else [])
    _compile(str(src), str(tmp_path / "t"), "c", defs)


def test_c_shim_rejects_long_double(tmp_path):
    src = tmp_path / "t.c"
    src.write_text("#define SDFT_FD_LONG_DOUBLE\n#include <sdft/sdft.h>\nint main(void){return 0;}\n")
    with pytest.raises(subprocess.CalledProcessError):
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include", "c"), str(src), "-o", str(tmp_path / "t")],
                       check=True, stderr=subprocess.DEVNULL)


def test_cpp_template_compiles_for_every_type_pair(tmp_path):
    src = tmp_path / "t.cpp"
    src.write_text("""
#include <sdft/sdft.h>
template <typename T, typename F> int use()
{
  try
  {
    sdft::SDFT<T, F> s(8, sdft::Window::Hamming, 0.5);
    T x[4] = {0}; std::complex<F> d[32]; std::complex<F>* rows[4] = {d, d + 8, d + 16, d + 24};
    const std::complex<F>* crows[4] = {d, d + 8, d + 16, d + 24};
    s.sdft(x[0], d); s.sdft(4, x, d); s.sdft(4, x, rows);
    x[0] = s.isdft(d); s.isdft(4, d, x); s.isdft(4, crows, x);
    s.reset();
    return (int)s.size() + (int)s.window() + (int)s.latency();
  }
  catch (const std::runtime_error&) { return -1; }
}
int main() { return (use<float, float>() + use<float, double>() + use<double, float>() + use<double, double>()) * 0; }
""")
    _compile(str(src), str(tmp_path / "t"), "cpp")


def test_drivers_compile():
    _compile(os.path.join(DRV, "hop_driver.c"), os.path.join(OUT, "hop_driver_c"), "c")
    _compile(os.path.join(DRV, "hop_driver_stream.c"), os.path.join(OUT, "hop_driver_stream_c"), "c")
    _compile(os.path.join(DRV, "hop_driver.cpp"), os.path.join(OUT, "hop_driver_cpp"), "cpp")


@pytest.mark.gpu
@pytest.mark.parametrize("lang", ["c", "cpp"])
def test_reference_shaped_drivers_match_oracle(lang, golden_dir):
    """test/main.sh parameters: DFTSIZE=1000 HOPSIZE=100 hann latency 1 on ALL 3528 hops of test.wav
    (test/main.py:67-79), the reference's own full integration test."""
    from oracle import Oracle
    exe = _compile(os.path.join(DRV, "hop_driver." + lang), os.path.join(OUT, "hop_driver_" + lang), lang)
    g = np.load(os.path.join(golden_dir, "testwav.npz"))
    x = (g["pcm24"].astype(np.float64) / 8388608.0).astype(np.float32)
    m, hop = 1000, 100
    assert x.size // hop == 3528
    res = subprocess.run([exe, str(m), str(hop), "1", "1"], input=x.tobytes(), stdout=subprocess.PIPE, check=True)
    nh = x.size // hop
    dfts = np.frombuffer(res.stdout[:nh * m * 16], np.complex128).reshape(nh, m)
    y = np.frombuffer(res.stdout[nh * m * 16:], np.float32)
    o = Oracle("f32", "f64", m, 1, 1.0)
    want_rows, want_y = [], []
    for h in range(nh):
        d = o.sdft(x[h * hop:(h + 1) * hop])
        want_rows.append(d[0].copy()); want_y.append(o.isdft(d))
    want_rows, want_y = np.stack(want_rows), np.concatenate(want_y)
    assert dfts.shape == want_rows.shape and y.shape == want_y.shape          # test/main.py:67-68, 75-76
    assert np.allclose(y, want_y)                                               # test/main.py:70
    assert np.allclose(dfts, want_rows)                                         # test/main.py:78
    assert np.abs(dfts - want_rows).max() / np.abs(want_rows).max() <= 1e-9     # BASELINE tolerance
    assert np.allclose(dfts[:24], g["t1000_rows"], rtol=0, atol=1e-12)          # golden rows generated from oracle/_ref


@pytest.mark.gpu
@pytest.mark.parametrize("depth", [1, 8])
def test_streaming_c_driver_on_device_buffers(depth, golden_dir):
    """The reference driver's hop loop (test/test.c:69-83) from plain C with every buffer on the device
    (sdft_b200_device_alloc, no CUDA toolkit on the caller's side) and the plan in streaming mode: m = 512,
    64-sample hops over the head of test.wav -- short calls below the 2m period, so every call rolls the history
    it got from its predecessor."""
    from oracle import Oracle
    exe = _compile(os.path.join(DRV, "hop_driver_stream.c"), os.path.join(OUT, "hop_driver_stream_c"), "c")
    g = np.load(os.path.join(golden_dir, "testwav.npz"))
    m, hop, nh = 512, 64, 1500
    x = (g["pcm24"].astype(np.float64) / 8388608.0).astype(np.float32)[50000:50000 + hop * nh]
    res = subprocess.run([exe, str(m), str(hop), "1", "1", str(depth)], input=x.tobytes(), stdout=subprocess.PIPE, check=True)
    dfts = np.frombuffer(res.stdout[:nh * m * 16], np.complex128).reshape(nh, m)
    y = np.frombuffer(res.stdout[nh * m * 16:], np.float32)
    o = Oracle("f32", "f64", m, 1, 1.0)
    want = o.sdft(x)
    want_y = o.isdft(want)
    assert np.abs(dfts - want[::hop]).max() / np.abs(want).max() <= 1e-9
    assert np.allclose(y, want_y, rtol=0, atol=2e-6)
