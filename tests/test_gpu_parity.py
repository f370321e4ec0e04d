"""
Parity of the CUDA path (through the C-ABI library) against the CPU oracle on identical inputs.
Tolerances are BASELINE.json's: max |X_gpu - X_ref| / max |X_ref| <= 1e-9 for double frequency-domain
data and <= 1e-4 for float; synthesized samples agree to float rounding.
"""
import ctypes
import itertools
import os

import numpy as np
import pytest

from conftest import seed_of

pytestmark = pytest.mark.gpu

TOL = {"f64": 1e-9, "f32": 1e-4}
TYPES = list(itertools.product(["f32", "f64"], repeat=2))


@pytest.fixture(scope="module")
def SDFT():
    import torch
    assert torch.cuda.is_available()
    from sdft_b200 import SDFT
    return SDFT


def rel_err(got, want):
    scale = np.abs(want).max()
    return np.abs(got - want).max() / (scale if scale > 0 else 1.0)


def _bits(a):
    return np.ascontiguousarray(a).view(np.uint8)


@pytest.mark.parametrize("fd", ["f32", "f64"])
@pytest.mark.parametrize("m", [1, 2, 3, 8, 37, 1000, 1024, 4096])
def test_tables_bit_identical(SDFT, fd, m):
    """Twiddle tables come from the host libm with the reference's expression order (sdft.h:439-446)."""
    from oracle import Oracle
    for latency in (1.0, 0.5, 0.3):
        g = SDFT(m, "hann", latency, td="f32", fd=fd)
        o = Oracle("f32", fd, m, "hann", latency)
        for a, b in zip(g.twiddles(), o.twiddles()):
            assert np.array_equal(_bits(a), _bits(b))


@pytest.mark.parametrize("td,fd", TYPES)
@pytest.mark.parametrize("window", ["boxcar", "hann", "hamming", "blackman"])
@pytest.mark.parametrize("latency", [1.0, 0.5])
def test_multicall_stream(SDFT, td, fd, window, latency):
    """Endless multi-call state: odd call sizes crossing the ring wrap and the modulation restart."""
    from oracle import Oracle
    rng = np.random.default_rng(seed_of(td, fd, window))
    for m in (1, 2, 3, 8, 37, 250, 1000):
        g = SDFT(m, window, latency, td=td, fd=fd)
        o = Oracle(td, fd, m, window, latency)
        for n in (1, 7, 100, 2 * m + 13, 5, 4 * m, 333):
            x = rng.uniform(-1, 1, n)
            want = o.sdft(x)
            got = g.sdft(x)
            assert got.shape == want.shape and got.dtype == want.dtype
            assert rel_err(got, want) <= TOL[fd], (m, n, rel_err(got, want))
            ys, yw = g.isdft(got), o.isdft(want)
            assert np.abs(ys.astype(np.float64) - yw.astype(np.float64)).max() <= (2e-6 if fd == "f64" else 2e-4), (m, n)
        cg, hg, ag, pg = g.state()
        co, ho, ao, po = o.state()
        assert cg == co
        assert np.array_equal(_bits(hg), _bits(ho))
        assert rel_err(ag, ao) <= TOL[fd]
        if fd == "f32":
            assert np.array_equal(_bits(pg), _bits(po)), "float modulation phase must be bit-exact"
        else:
            assert np.abs(pg - po).max() <= 1e-12


@pytest.mark.parametrize("fd", ["f32", "f64"])
@pytest.mark.parametrize("chunk", [32, 64, 96, 256, 1024])
def test_chunk_lengths(SDFT, fd, chunk):
    """Any chunk length (multiple of 32) gives the same rows; resets fall inside and between chunks."""
    from oracle import Oracle
    rng = np.random.default_rng(chunk)
    for m in (37, 1000, 1024):
        g = SDFT(m, "blackman", 0.5, td="f32", fd=fd)
        g.set_chunk(chunk)
        o = Oracle("f32", fd, m, "blackman", 0.5)
        for n in (777, 2 * m + 13, 3000):
            x = rng.uniform(-1, 1, n).astype(np.float32)
            want, got = o.sdft(x), g.sdft(x)
            assert rel_err(got, want) <= TOL[fd], (m, n, rel_err(got, want))


@pytest.mark.parametrize("window", ["boxcar", "hann", "hamming", "blackman"])
def test_double_modulated_mode(SDFT, window, monkeypatch):
    """SDFT_B200_F64=modulated runs the reference's modulated scheme in double as well (the default for
    double is the demodulated fast replay); both must sit inside the 1e-9 gate, and close to each other."""
    from oracle import Oracle
    rng = np.random.default_rng(77)
    for m in (2, 37, 1000):
        fast = SDFT(m, window, 0.5, td="f32", fd="f64")
        monkeypatch.setenv("SDFT_B200_F64", "modulated")
        exact = SDFT(m, window, 0.5, td="f32", fd="f64")
        monkeypatch.delenv("SDFT_B200_F64")
        o = Oracle("f32", "f64", m, window, 0.5)
        for n in (7, 2 * m + 13, 3000, 1):
            x = rng.uniform(-1, 1, n).astype(np.float32)
            want, a, b = o.sdft(x), fast.sdft(x), exact.sdft(x)
            assert rel_err(b, want) <= 1e-12, (m, n, rel_err(b, want))
            assert rel_err(a, want) <= 1e-9, (m, n, rel_err(a, want))
            assert rel_err(a, b) <= 1e-10
        assert rel_err(fast.state()[2], o.state()[2]) <= 1e-9
        assert rel_err(exact.state()[2], o.state()[2]) <= 1e-12
        # the fused round trip and the state-only pass in the literal mode as well
        x = rng.uniform(-1, 1, 2 * m + 50).astype(np.float32)
        want_y = o.roundtrip(x)
        assert np.abs(exact.roundtrip(x) - want_y).max() <= 2e-6
        x2 = rng.uniform(-1, 1, 90).astype(np.float32)
        exact.advance(x2)
        o.advance(x2)
        assert rel_err(exact.state()[2], o.state()[2]) <= 1e-12


def test_float_rows_bit_exact_within_a_chunk(SDFT):
    """The float path keeps every rounding of the reference: with one chunk per call it follows the
    reference's summation order exactly and the rows are bit-identical."""
    from oracle import Oracle
    m = 64
    g = SDFT(m, "hann", 1, td="f32", fd="f32")
    g.set_chunk(1024)
    o = Oracle("f32", "f32", m, "hann", 1.0)
    x = np.random.default_rng(5).uniform(-1, 1, 2 * m).astype(np.float32)
    want, got = o.sdft(x), g.sdft(x)
    assert np.array_equal(_bits(got), _bits(want))


@pytest.mark.parametrize("window", ["boxcar", "hann", "hamming", "blackman"])
def test_float_fast_totals_vs_strict(SDFT, window, monkeypatch):
    """Default float mode: the chunk totals that feed the carries are summed in double (FP64 pipe), the
    replay keeps every rounding of the reference.  SDFT_B200_F32=strict sums the totals with the float
    recurrence as well.  Both sit far inside the 1e-4 gate and within a few 1e-6 of each other; the
    modulation phase is bit-exact in both."""
    from oracle import Oracle
    rng = np.random.default_rng(78)
    for m in (3, 37, 1000):
        fast = SDFT(m, window, 0.5, td="f32", fd="f32")
        monkeypatch.setenv("SDFT_B200_F32", "strict")
        strict = SDFT(m, window, 0.5, td="f32", fd="f32")
        monkeypatch.delenv("SDFT_B200_F32")
        o = Oracle("f32", "f32", m, window, 0.5)
        for n in (7, 2 * m + 13, 3000, 1, 700):
            x = rng.uniform(-1, 1, n).astype(np.float32)
            want, a, b = o.sdft(x), fast.sdft(x), strict.sdft(x)
            assert rel_err(b, want) <= 1e-5, (m, n, rel_err(b, want))
            assert rel_err(a, want) <= 1e-5, (m, n, rel_err(a, want))
            assert rel_err(a, b) <= 1e-5
        assert np.array_equal(_bits(fast.state()[3]), _bits(o.state()[3])), "phase must stay bit-exact"
        assert rel_err(fast.state()[2], o.state()[2]) <= 1e-5


@pytest.mark.parametrize("seed", range(8))
def test_randomized_plans_and_call_sequences(SDFT, seed, monkeypatch):
    """Random plan shapes, launch geometries and call sequences against the oracle: exercises chunk
    boundaries, period wraps inside and between calls, partial first chunks, both warp geometries, every
    CTA width and chunk length, single- and multi-block chains."""
    from oracle import Oracle
    rng = np.random.default_rng(1000 + seed)
    for trial in range(6):
        td, fd = TYPES[int(rng.integers(4))]
        window = ["boxcar", "hann", "hamming", "blackman"][int(rng.integers(4))]
        latency = float(rng.choice([1.0, 0.5, 0.25]))
        m = int(rng.choice([1, 2, 3, 5, 8, 31, 64, 100, 257, 600, 1000]))
        monkeypatch.setenv("SDFT_B200_GEO", str(rng.choice(["wide", "narrow"])))
        monkeypatch.setenv("SDFT_B200_WARPS", str(int(rng.choice([0, 1, 2, 3, 4, 8]))))
        g = SDFT(m, window, latency, td=td, fd=fd)
        chunk = int(rng.choice([0, 32, 64, 96, 128, 512]))
        if chunk:
            g.set_chunk(chunk)
        o = Oracle(td, fd, m, window, latency)
        for call in range(6):
            kind = int(rng.integers(4))
            n = [int(rng.integers(1, 10)), int(rng.integers(1, 2 * m + 2)), 2 * m + int(rng.integers(-3, 4)),
                 int(rng.integers(2 * m, 6 * m + 50))][kind]
            n = max(1, n)
            x = rng.uniform(-1, 1, n)
            want, got = o.sdft(x), g.sdft(x)
            assert rel_err(got, want) <= TOL[fd], (seed, trial, call, td, fd, window, m, n, chunk, rel_err(got, want))
        cg, hg, ag, _ = g.state()
        co, ho, ao, _ = o.state()
        assert cg == co and np.array_equal(_bits(hg), _bits(ho)) and rel_err(ag, ao) <= TOL[fd]


@pytest.mark.parametrize("m,fd", [(16384, "f64"), (10007, "f64"), (10007, "f32"), (8192, "f32"), (65536, "f64"), (65536, "f32")])
def test_large_and_odd_dft_sizes(SDFT, m, fd):
    """Big plans and an odd prime size (rows not 32-byte aligned: the bin-by-bin store path), over calls that
    cross the 2m period.  The reference's plan is O(m) (sdft.h:428-437); ours keeps its phase source O(m) for
    double (2m roots of unity) and within a fixed budget for float (table stride growing with m)."""
    from oracle import Oracle
    rng = np.random.default_rng(m)
    g = SDFT(m, "blackman", 0.5, td="f32", fd=fd)
    assert g._lib.sdft_b200_table_bytes(g._h) < 256 << 20
    if fd == "f64":
        assert g._lib.sdft_b200_table_bytes(g._h) <= 100 * m          # O(m): 16 B x (m+4 + m + m+4 + 2m)
    o = Oracle("f32", fd, m, "blackman", 0.5)
    if m >= 32768 and fd == "f32":
        # 700 samples into a 131072-sample Blackman window every row is the float cancellation noise of taps
        # that sum to ~0 (|row| ~ 1e-9 |accumulator|): the reference's rows are its own rounding noise there and
        # nothing can match them to 1e-4 of their maximum -- except the same roundings in the same order.  So: one
        # chunk per call with the float recurrence in the totals too (the literal mode) must be BIT-exact, which
        # pins the strided phase table; the tolerance comparison follows once the window has filled.
        import os
        import torch
        os.environ["SDFT_B200_F32"] = "strict"
        try:
            lit = SDFT(m, "blackman", 0.5, td="f32", fd=fd)
        finally:
            del os.environ["SDFT_B200_F32"]
        lit.set_chunk(1024)
        o2 = Oracle("f32", fd, m, "blackman", 0.5)
        # device rows: one launch, one chunk (host rows would be produced in 128 MiB tiles = several calls)
        x0 = rng.uniform(-1, 1, 700).astype(np.float32)
        assert np.array_equal(_bits(lit.sdft(torch.from_numpy(x0).cuda()).cpu().numpy()), _bits(o2.sdft(x0)))
        lit.reset()
        o2.reset()
        x1 = rng.uniform(-1, 1, 1000).astype(np.float32)
        lit.advance(x1[:700])                                    # the next call starts at cursor 700: off the table grid
        o2.advance(x1[:700])
        got1 = lit.sdft(torch.from_numpy(x1[700:]).cuda()).cpu().numpy()
        want1 = o2.sdft(x1[700:])
        assert rel_err(got1, want1) <= 1e-6                      # the carry of the state-only call is summed in its own order
        assert np.array_equal(_bits(lit.state()[3]), _bits(o2.state()[3])), "phase after 1000 samples, table stride 512"
        del lit, o2
    else:
        for n in (700, 1500, 1):
            x = rng.uniform(-1, 1, n).astype(np.float32)
            want, got = o.sdft(x), g.sdft(x)
            assert rel_err(got, want) <= TOL[fd], (m, n, rel_err(got, want))
            assert np.abs(g.isdft(got).astype(np.float64) - o.isdft(want)).max() <= (2e-6 if fd == "f64" else 2e-4)
    x2 = rng.uniform(-1, 1, 2 * m).astype(np.float32)            # across the period boundary, state only
    g.advance(x2)
    o.advance(x2)
    x3 = rng.uniform(-1, 1, 40).astype(np.float32)
    assert rel_err(g.sdft(x3), o.sdft(x3)) <= TOL[fd]
    assert g.state()[0] == o.state()[0]


@pytest.mark.parametrize("m,window", [(2048, "hann"), (1024, "blackman"), (2000, "hamming")])
def test_float_calls_split_into_wide_body_and_narrow_tail(SDFT, m, window, monkeypatch):
    """Long float calls whose last wide warp group would be mostly empty run as two chain sets over disjoint bin
    ranges in one launch (sdft_launch.hpp: ScanPart, scan_emit_mixed_kernel): rows, synthesized samples and state against the oracle over several calls,
    with a state-only call and a region of interest in between, and against the same plan with the split
    disabled (identical up to the order in which the tail's carries are added)."""
    import torch
    from oracle import Oracle
    n = 1700000 // (-(-m // 248)) + 1000          # enough warp-steps for the wide geometry (choose_geo), >= 2^26 bin-updates
    rng = np.random.default_rng(seed_of("split", m, window))
    g = SDFT(m, window, 0.5, td="f32", fd="f32")
    monkeypatch.setenv("SDFT_B200_NO_SPLIT", "1")
    one = SDFT(m, window, 0.5, td="f32", fd="f32")
    monkeypatch.delenv("SDFT_B200_NO_SPLIT")
    o = Oracle("f32", "f32", m, window, 0.5)
    launches0 = g.launches
    for call in range(3):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        xt = torch.from_numpy(x).cuda()
        if call == 1:
            g.advance(xt); one.advance(xt); o.advance(x)
            continue
        got, ref = g.sdft(xt), one.sdft(xt)
        want = o.sdft(x)
        assert rel_err(got.cpu().numpy(), want) <= TOL["f32"], (call, rel_err(got.cpu().numpy(), want))
        assert float((got - ref).abs().max()) <= 1e-5 * float(ref.abs().max())
        y, yw = g.isdft(got).cpu().numpy(), o.isdft(want)
        assert np.abs(y.astype(np.float64) - yw).max() <= 2e-4 * max(np.abs(yw).max(), 1e-3)
    assert g.launches - launches0 == 3 + 2, "one scan launch per call (+ two synthesis launches)"
    assert g._lib.sdft_b200_split_count(g._h) == 3 and one._lib.sdft_b200_split_count(one._h) == 0
    cg, hg, ag, pg = g.state()
    co, ho, ao, po = o.state()
    assert cg == co and np.array_equal(_bits(hg), _bits(ho)) and np.array_equal(_bits(pg), _bits(po))
    assert rel_err(ag, ao) <= TOL["f32"]
    # the last bins (the tail launch's) through a region of interest
    g.set_roi(m - 96, 96)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    got = g.sdft(torch.from_numpy(x).cuda()).cpu().numpy()
    want = o.sdft(x)
    assert got.shape == (n, 96)
    assert np.abs(got - want[:, m - 96:]).max() <= TOL["f32"] * np.abs(want).max()


@pytest.mark.parametrize("budget_mb,m", [(1, 4096), (0, 1000), (2, 10007)])
def test_float_phase_table_with_a_coarse_stride(SDFT, budget_mb, m, monkeypatch):
    """The float phase table holds the reference's sequential fiddle recurrence at every `stride`-th cursor
    and rotates the remainder; a small budget forces strides of 64..1024 on ordinary sizes.  The phase must
    stay BIT-exact at any cursor and the rows inside the float gate, for calls starting anywhere in the period."""
    from oracle import Oracle
    monkeypatch.setenv("SDFT_B200_F0_BUDGET_MB", str(budget_mb))
    coarse = SDFT(m, "hamming", 0.5, td="f32", fd="f32")
    monkeypatch.delenv("SDFT_B200_F0_BUDGET_MB")
    fine = SDFT(m, "hamming", 0.5, td="f32", fd="f32")
    assert coarse._lib.sdft_b200_table_bytes(coarse._h) < fine._lib.sdft_b200_table_bytes(fine._h)
    o = Oracle("f32", "f32", m, "hamming", 0.5)
    rng = np.random.default_rng(seed_of(budget_mb, m))
    for n in (77, 1500, 2 * m + 333, 1, 900, 31):
        x = rng.uniform(-1, 1, n).astype(np.float32)
        want = o.sdft(x)
        got_c, got_f = coarse.sdft(x), fine.sdft(x)
        assert rel_err(got_c, want) <= TOL["f32"], (m, n, rel_err(got_c, want))
        assert rel_err(got_c, got_f) <= 1e-5
        pc, po = coarse.state()[3], o.state()[3]
        assert np.array_equal(_bits(pc), _bits(po)), "float modulation phase must be bit-exact at any table stride"


def test_oversized_float_plan_is_rejected_with_a_message(SDFT):
    """A float frequency-domain plan whose phase table cannot fit fails in sdft_alloc with an explanation, not
    with a cudaMalloc error (the double frequency domain has no such limit)."""
    with pytest.raises(RuntimeError, match="phase table"):
        SDFT(1 << 22, "hann", 1, td="f32", fd="f32")


def test_reset_and_getters(SDFT):
    from oracle import Oracle
    g = SDFT(16, "hamming", 0.5, td="f32", fd="f64")
    x = np.linspace(-1, 1, 77).astype(np.float32)
    first = g.sdft(x)
    g.reset()
    again = g.sdft(x)
    assert np.array_equal(_bits(first), _bits(again))
    lib = g._lib
    assert lib.sdft_b200_f32f64_size(g._h) == 16
    assert lib.sdft_b200_f32f64_window(g._h) == 2
    assert lib.sdft_b200_f32f64_latency(g._h) == 0.5
    # NULL handling of sdft.h:468, 537, 545, 553
    assert lib.sdft_b200_f32f64_size(None) == 0
    assert lib.sdft_b200_f32f64_window(None) == 0
    assert lib.sdft_b200_f32f64_latency(None) == 0.0
    lib.sdft_b200_f32f64_free(None)
    # default plan = hann, latency 1 (sdft.h:457-460)
    h = ctypes.c_void_p(lib.sdft_b200_f32f64_alloc(8))
    assert lib.sdft_b200_f32f64_window(h) == 1 and lib.sdft_b200_f32f64_latency(h) == 1.0
    lib.sdft_b200_f32f64_free(h)


def test_single_sample_and_row_pointer_variants(SDFT):
    """sdft_sdft / sdft_isdft (sdft.h:562, 635) and the _nd variants (sdft.h:622, 681)."""
    from oracle import Oracle
    m, n = 24, 50
    g = SDFT(m, "hann", 1, td="f32", fd="f64")
    o = Oracle("f32", "f64", m, "hann", 1.0)
    lib = g._lib
    x = np.random.default_rng(11).uniform(-1, 1, n).astype(np.float32)
    want = o.sdft(x)
    row = np.empty(m, np.complex128)
    for t in range(10):
        lib.sdft_b200_f32f64_sdft(g._h, ctypes.c_float(x[t]), row.ctypes.data_as(ctypes.c_void_p))
        assert rel_err(row, want[t]) <= 1e-9
        y = lib.sdft_b200_f32f64_isdft(g._h, row.ctypes.data_as(ctypes.c_void_p))
        assert abs(y - o.isdft(want[t:t + 1])[0]) <= 2e-6
    rows = [np.empty(m, np.complex128) for _ in range(n - 10)]
    ptrs = (ctypes.c_void_p * len(rows))(*[r.ctypes.data for r in rows])
    xs = np.ascontiguousarray(x[10:])
    lib.sdft_b200_f32f64_sdft_nd(g._h, len(rows), xs.ctypes.data_as(ctypes.c_void_p), ptrs)
    got = np.stack(rows)
    assert rel_err(got, want[10:]) <= 1e-9
    y = np.empty(len(rows), np.float32)
    lib.sdft_b200_f32f64_isdft_nd(g._h, len(rows), ptrs, y.ctypes.data_as(ctypes.c_void_p))
    assert np.abs(y - o.isdft(want[10:])).max() <= 2e-6


@pytest.mark.parametrize("layout", ["contiguous", "runs", "scattered_host", "scattered_device", "mixed"])
def test_row_pointer_variants_at_size(SDFT, layout):
    """sdft_sdft_nd / sdft_isdft_nd (sdft.h:622-628, :681-687) at n = 2^16, m = 1024 for every way a caller can
    lay its rows out: slices of one matrix (one run), a ring of hop buffers (a few long runs), rows scattered
    through host memory, through device memory, and both kinds in one call.  Against sdft_sdft_n / sdft_isdft_n
    of a twin plan (bit-identical: same kernels, only the destination differs) and the oracle on a sample."""
    import torch
    from oracle import Oracle
    m, n = 1024, 1 << 16
    x = np.random.default_rng(seed_of("nd", layout)).uniform(-1, 1, n).astype(np.float32)
    twin = SDFT(m, "hann", 0.5, td="f32", fd="f64")
    want = twin.sdft(x)
    want_y = twin.isdft(want)
    g = SDFT(m, "hann", 0.5, td="f32", fd="f64")
    lib = g._lib
    perm = np.random.default_rng(5).permutation(n)
    host = np.zeros((n, m), np.complex128)
    dev = torch.zeros((n, m), dtype=torch.complex128, device="cuda")
    if layout == "contiguous":
        addr = [host.ctypes.data + i * m * 16 for i in range(n)]
    elif layout == "runs":                       # 16 hop buffers of 4096 rows, in shuffled order
        order = np.random.default_rng(6).permutation(16)
        addr = [host.ctypes.data + (int(order[i // 4096]) * 4096 + i % 4096) * m * 16 for i in range(n)]
    elif layout == "scattered_host":
        addr = [host.ctypes.data + int(perm[i]) * m * 16 for i in range(n)]
    elif layout == "scattered_device":
        addr = [dev.data_ptr() + int(perm[i]) * m * 16 for i in range(n)]
    else:
        addr = [(dev.data_ptr() if perm[i] % 2 else host.ctypes.data) + int(perm[i]) * m * 16 for i in range(n)]
    ptrs = (ctypes.c_void_p * n)(*addr)
    lib.sdft_b200_f32f64_sdft_nd(g._h, n, x.ctypes.data_as(ctypes.c_void_p), ptrs)
    g.synchronize()
    dev_h = dev.cpu().numpy()
    base_h, base_d = host.ctypes.data, dev.data_ptr()
    idx = np.array([((a - base_d) if (layout in ("scattered_device", "mixed") and base_d <= a < base_d + n * m * 16) else (a - base_h)) // (m * 16)
                    for a in addr])
    on_dev = np.array([layout in ("scattered_device", "mixed") and base_d <= a < base_d + n * m * 16 for a in addr])
    got = np.where(on_dev[:, None], dev_h[idx], host[idx])
    assert np.array_equal(_bits(got), _bits(want))
    y = np.zeros(n, np.float32)
    lib.sdft_b200_f32f64_isdft_nd(g._h, n, ptrs, y.ctypes.data_as(ctypes.c_void_p))
    g._check()
    assert np.array_equal(_bits(y), _bits(want_y))
    o = Oracle("f32", "f64", m, "hann", 0.5)
    ref = o.sdft(x[:3000])
    assert rel_err(got[:3000], ref) <= 1e-9


def test_device_pointers_and_batch(SDFT):
    """Device-resident path (torch tensors) and the batched multi-channel plan."""
    import torch
    from oracle import Oracle
    m, n, ch = 250, 3000, 3
    rng = np.random.default_rng(21)
    x = rng.uniform(-1, 1, (ch, n)).astype(np.float32)
    g = SDFT(m, "hann", 0.5, td="f32", fd="f64", channels=ch)
    xt = torch.from_numpy(x).cuda()
    parts, ys = [], []
    for a, b in ((0, 1000), (1000, 1001), (1001, 3000)):
        d = g.sdft(xt[:, a:b].contiguous())
        parts.append(d)
        ys.append(g.isdft(d))
    torch.cuda.synchronize()
    got = torch.cat(parts, dim=1).cpu().numpy()
    y = torch.cat(ys, dim=1).cpu().numpy()
    for c in range(ch):
        o = Oracle("f32", "f64", m, "hann", 0.5)
        want = o.sdft(x[c])
        assert rel_err(got[c], want) <= 1e-9, c
        assert np.abs(y[c] - o.isdft(want)).max() <= 2e-6
    # host batch path
    g2 = SDFT(m, "hann", 0.5, td="f32", fd="f64", channels=ch)
    got2 = g2.sdft(x)
    assert rel_err(got2, got) <= 1e-13   # one 3000-sample call vs three calls: other chunking, same rows


@pytest.mark.parametrize("fd", ["f32", "f64"])
def test_convolve_matches_reference_python_class(SDFT, fd, golden_dir):
    """SDFT.convolve (python/src/sdft/sdft.py:146-203) against vectors produced by the reference's own
    Python class (tests/golden/make_golden.py), host and device inputs."""
    import torch
    g = np.load(os.path.join(golden_dir, "py_convolve.npz"))
    tol = 1e-13 if fd == "f64" else 1e-6
    for m in (8, 37):
        x = g["x_m%d" % m]
        for window in ("boxcar", "hann", "hamming", "blackman"):
            want = g["y_m%d_%s" % (m, window)]
            plan = SDFT(m, window, 1, td="f32", fd=fd)
            got = plan.convolve(x)
            assert got.shape == want.shape
            assert np.abs(got - want).max() <= tol * np.abs(want).max(), (m, window)
            got_dev = plan.convolve(torch.from_numpy(x).cuda()).cpu().numpy()
            assert np.array_equal(_bits(got_dev), _bits(got))


@pytest.mark.parametrize("td,fd", [("f32", "f64"), ("f32", "f32")])
@pytest.mark.parametrize("window", ["hann", "blackman"])
def test_region_of_interest_rows(SDFT, td, fd, window, monkeypatch):
    """sdft_b200_set_roi: rows of a sub-band are the corresponding columns of the full rows, bit for bit
    (group-aligned and unaligned regions, both warp geometries, host and device outputs), the plan state is
    unaffected, and isdft of such rows is the band-limited sum over the region."""
    import torch
    from oracle import Oracle
    rng = np.random.default_rng(61)
    m = 300
    x = rng.uniform(-1, 1, 2 * m + 77).astype(np.float32)
    for geo in ("wide", "narrow"):
        monkeypatch.setenv("SDFT_B200_GEO", geo)
        full = SDFT(m, window, 0.5, td=td, fd=fd)
        want = np.concatenate([full.sdft(x[:200]), full.sdft(x[200:])])      # same call split as below
        for first, count in ((0, 64), (8, 120), (37, 101), (m - 5, 5), (123, 1), (0, m)):
            sub = SDFT(m, window, 0.5, td=td, fd=fd)
            sub.set_roi(first, count)
            got = sub.sdft(x[:200])
            got2 = sub.sdft(torch.from_numpy(x[200:]).cuda()).cpu().numpy()
            got = np.concatenate([got, got2])
            assert got.shape == (x.size, count)
            assert np.array_equal(_bits(got), _bits(np.ascontiguousarray(want[:, first:first + count]))), (geo, first, count)
            for a, b in zip(sub.state()[1:3], full.state()[1:3]):
                assert np.array_equal(_bits(a), _bits(b))
            # band-limited synthesis: the oracle's isdft of rows that are zero outside the region
            o = Oracle(td, fd, m, window, 0.5)
            masked = np.zeros_like(want)
            masked[:, first:first + count] = want[:, first:first + count]
            y_want = o.isdft(masked).astype(np.float64)
            y_got = sub.isdft(got).astype(np.float64)
            assert np.abs(y_got - y_want).max() <= (2e-6 if fd == "f64" else 2e-4) * max(1.0, np.abs(y_want).max())


def test_cuda_array_interface_inputs(SDFT):
    """Anything that exposes __cuda_array_interface__ is taken zero-copy (CuPy, Numba, ...)."""
    import torch

    class Foreign:
        def __init__(self, t):
            self._t = t
            self.__cuda_array_interface__ = t.__cuda_array_interface__

    x = np.random.default_rng(3).uniform(-1, 1, 500).astype(np.float32)
    a, b = SDFT(32, "hann", 1, td="f32", fd="f64"), SDFT(32, "hann", 1, td="f32", fd="f64")
    xt = torch.from_numpy(x).cuda()
    got = a.sdft(Foreign(xt))
    want = b.sdft(xt)
    assert torch.equal(got, want)
    assert torch.equal(a.isdft(Foreign(got)), b.isdft(want))


def test_advance_and_roundtrip(SDFT):
    from oracle import Oracle
    m = 128
    rng = np.random.default_rng(31)
    x = rng.uniform(-1, 1, 4 * m + 100).astype(np.float32)
    o = Oracle("f32", "f64", m, "hann", 1.0)
    want = o.sdft(x)
    g = SDFT(m, "hann", 1, td="f32", fd="f64")
    g.advance(x[:2 * m + 7])
    got = g.sdft(x[2 * m + 7:])
    assert rel_err(got, want[2 * m + 7:]) <= 1e-9
    g.reset()
    y = g.roundtrip(x)
    assert np.abs(y - o.isdft(want)).max() <= 2e-6


@pytest.mark.parametrize("td,fd", TYPES)
@pytest.mark.parametrize("window", ["boxcar", "hann", "hamming", "blackman"])
@pytest.mark.parametrize("latency", [1.0, 0.5])
def test_fused_roundtrip(SDFT, td, fd, window, latency):
    """sdft_b200_*_roundtrip_n: analysis and synthesis in one kernel (the rows never exist) against the
    reference's sdft_sdft followed by sdft_isdft, sample by sample (test/test.c:79-80), over several
    calls on one plan; the plan state must end up exactly where the row-producing path leaves it."""
    from oracle import Oracle
    rng = np.random.default_rng(seed_of(td, fd, window, latency))
    tol = {("f32", "f64"): 2e-6, ("f64", "f64"): 1e-9, ("f32", "f32"): 2e-4, ("f64", "f32"): 2e-4}[(td, fd)]
    for m in (3, 37, 250, 1000):
        g = SDFT(m, window, latency, td=td, fd=fd)
        rows = SDFT(m, window, latency, td=td, fd=fd)
        o = Oracle(td, fd, m, window, latency)
        for n in (1, 9, 2 * m + 13, 4 * m, 333, 5000):
            x = rng.uniform(-1, 1, n)
            want = o.roundtrip(x).astype(np.float64)
            got = g.roundtrip(x).astype(np.float64)
            scale = max(np.abs(want).max(), 1e-3)
            assert np.abs(got - want).max() <= tol * scale, (m, n, np.abs(got - want).max() / scale)
            rows.sdft(x)
        cg, hg, ag, _ = g.state()
        cr, hr, ar, _ = rows.state()
        assert cg == cr and np.array_equal(_bits(hg), _bits(hr)) and np.array_equal(_bits(ag), _bits(ar))


@pytest.mark.parametrize("geo", ["wide", "narrow", None])
@pytest.mark.parametrize("fd", ["f32", "f64"])
@pytest.mark.parametrize("window", ["hann", "hamming", "blackman"])
def test_rows_after_fused_roundtrip_on_one_plan(SDFT, window, fd, geo, monkeypatch):
    """A fused round trip runs the halo-free geometry, where no warp owns the mirror cells of
    sdft.h:589-595; the windowed row kernel of the NEXT call on the same plan reads them as its carry
    (bins 0-1 and m-2, m-1 depend on them).  roundtrip -> sdft -> advance -> roundtrip -> sdft on one
    plan against the oracle, with m a multiple of the warp width (cells m+2, m+3 unowned too) and not."""
    from oracle import Oracle
    if geo:
        monkeypatch.setenv("SDFT_B200_GEO", geo)
    rng = np.random.default_rng(seed_of(window, fd, geo))
    for m in (128, 256, 512, 250, 1000, 3, 2, 1):
        g = SDFT(m, window, 0.5, td="f32", fd=fd)
        o = Oracle("f32", fd, m, window, 0.5)
        for n_rt, n_rows in ((2 * m + 50, 333), (7, 2 * m + 13), (700, 64)):
            x1 = rng.uniform(-1, 1, n_rt).astype(np.float32)
            x2 = rng.uniform(-1, 1, n_rows).astype(np.float32)
            want_y, got_y = o.roundtrip(x1).astype(np.float64), g.roundtrip(x1).astype(np.float64)
            assert np.abs(got_y - want_y).max() <= (2e-6 if fd == "f64" else 2e-4) * max(np.abs(want_y).max(), 1e-3)
            want, got = o.sdft(x2), g.sdft(x2)
            assert rel_err(got, want) <= TOL[fd], (m, n_rt, n_rows, rel_err(got, want))
            # the quirk bins and bin 0 on their own: they are the ones the mirror cells feed
            for k in {0, min(1, m - 1), max(m - 2, 0), m - 1}:
                assert np.abs(got[:, k] - want[:, k]).max() <= TOL[fd] * np.abs(want).max(), (m, k)
            x3 = rng.uniform(-1, 1, 90).astype(np.float32)
            g.advance(x3)
            o.advance(x3)


@pytest.mark.parametrize("latency", [1.0, 0.5])
@pytest.mark.parametrize("fd", ["f32", "f64"])
def test_fused_roundtrip_with_spectral_gain(SDFT, fd, latency):
    """roundtrip_gain_n: isdft(gains .* sdft(x)) without the matrix; against the oracle's rows scaled on
    the host and synthesized by the oracle."""
    from oracle import Oracle
    m, n = 250, 3000
    rng = np.random.default_rng(29)
    x = rng.uniform(-1, 1, n).astype(np.float32)
    gains = (rng.uniform(0, 1, m) * np.exp(2j * np.pi * rng.uniform(0, 1, m))).astype(np.complex128)
    o = Oracle("f32", fd, m, "hann", latency)
    want = o.isdft((o.sdft(x).astype(np.complex128) * gains).astype(o.fdx_np)).astype(np.float64)
    g = SDFT(m, "hann", latency, td="f32", fd=fd)
    got = np.concatenate([g.roundtrip(x[:1700], gains=gains), g.roundtrip(x[1700:], gains=gains)]).astype(np.float64)
    tol = 2e-6 if fd == "f64" else 2e-4
    assert np.abs(got - want).max() <= tol * max(np.abs(want).max(), 1e-3)


def test_fused_roundtrip_batch_and_pieces(SDFT, monkeypatch):
    """Batched channels, device tensors, and a call cut into several scratch-bounded pieces."""
    import torch
    from oracle import Oracle
    m, n, ch = 250, 6000, 3
    x = np.random.default_rng(23).uniform(-1, 1, (ch, n)).astype(np.float32)
    monkeypatch.setenv("SDFT_B200_ROUNDTRIP_PIECE", "1024")
    g = SDFT(m, "hamming", 0.5, td="f32", fd="f64", channels=ch)
    y = g.roundtrip(torch.from_numpy(x).cuda()).cpu().numpy()
    y2 = SDFT(m, "hamming", 0.5, td="f32", fd="f64", channels=ch).roundtrip(x)
    for c in range(ch):
        want = Oracle("f32", "f64", m, "hamming", 0.5).roundtrip(x[c])
        assert np.abs(y[c] - want).max() <= 2e-6
        assert np.abs(y2[c] - want).max() <= 2e-6


@pytest.mark.parametrize("fd,depth", [("f64", 4), ("f64", 8), ("f32", 3)])
def test_streaming_mode_endless_hops(SDFT, fd, depth):
    """Streaming mode (sdft_b200_set_streaming): 2^20 samples in 256 calls of 4096 on one plan, m = 512 -- the
    hop loop of test/test.c:69-83 with device buffers (BASELINE config 5's shape), consecutive calls overlapping
    on the GPU.  With the chunk length pinned the rows are BIT-identical to the serial mode (same kernels, same
    summation order, only the waiting differs); with the streaming default (longer chunks) they differ by the
    order in which the carries are added.  Against the oracle on sampled calls; state identical at the end."""
    import torch
    from oracle import Oracle
    m, hop, calls = 512, 4096, 256
    n = hop * calls
    x = np.random.default_rng(seed_of("stream", fd, depth)).uniform(-1, 1, n).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    cdt = torch.complex128 if fd == "f64" else torch.complex64

    def run(plan_depth, chunk, hops):
        g = SDFT(m, "hann", 1, td="f32", fd=fd)
        g.set_streaming(plan_depth)
        if chunk:
            g.set_chunk(chunk)
        out = torch.zeros((calls, hop, m), dtype=cdt, device="cuda")
        if hops:
            g.sdft_hops(xt, hop, out)            # the same hop loop issued from inside the library
        else:
            for c in range(calls):
                g.sdft(xt[c * hop:(c + 1) * hop], out=out[c])
        g.synchronize()
        return g, torch.view_as_real(out)

    serial, want = run(1, 64, False)
    scoped = SDFT(m, "hann", 1, td="f32", fd=fd)
    scoped.set_chunk(64)
    with scoped.streaming(depth):                 # the context-manager spelling; leaves the plan serial again
        first = torch.view_as_real(scoped.sdft(xt[:hop]))
    assert torch.equal(first, want[0])
    flow, got = run(depth, 64, False)
    assert torch.equal(got, want)
    for a, b in zip(flow.state(), serial.state()):
        assert np.array_equal(_bits(np.asarray(a)), _bits(np.asarray(b)))
    _, got_hops = run(depth, 64, True)
    assert torch.equal(got_hops, want)
    del got_hops
    auto, got_auto = run(depth, 0, True)         # the streaming default: longer chunks than a serial call picks
    scale = float(want.abs().max())
    assert float((got_auto - want).abs().max()) <= (1e-12 if fd == "f64" else 2e-5) * scale
    _, again = run(depth, 0, False)
    assert torch.equal(again, got_auto), "streaming results must not depend on timing"
    del again, want
    o = Oracle("f32", fd, m, "hann", 1.0)
    rows = torch.view_as_complex(got_auto)
    for c in range(calls):
        if c in (0, 1, 2, 17, 100, calls - 1):
            ref = o.sdft(x[c * hop:(c + 1) * hop])
            assert rel_err(rows[c].cpu().numpy(), ref) <= TOL[fd], c
            assert rel_err(torch.view_as_complex(got)[c].cpu().numpy(), ref) <= TOL[fd], c
        else:
            o.advance(x[c * hop:(c + 1) * hop])
    assert rel_err(auto.state()[2], o.state()[2]) <= TOL[fd]
    assert rel_err(flow.state()[2], o.state()[2]) <= TOL[fd]


@pytest.mark.parametrize("td,fd", [("f32", "f64"), ("f32", "f32"), ("f64", "f64")])
@pytest.mark.parametrize("window", ["boxcar", "blackman"])
def test_streaming_mode_mixed_call_sequences(SDFT, td, fd, window):
    """Streaming with everything the state hand-over has to survive: call sizes below and above the 2m period
    (short calls roll the history from the previous call's), state-only calls, fused round trips and host-buffer
    calls (both serial by construction) in between, a reset, a change of depth in mid-stream."""
    import torch
    from oracle import Oracle
    rng = np.random.default_rng(seed_of("mixed", td, fd, window))
    tdt = torch.float32 if td == "f32" else torch.float64
    for m in (37, 250, 1024):
        g = SDFT(m, window, 0.5, td=td, fd=fd)
        g.set_streaming(5)
        o = Oracle(td, fd, m, window, 0.5)
        sizes = [1, 7, 100, 2 * m + 13, 5, 4 * m, 333, 64, 64, 64, 3000, 2, 900]
        for rnd in range(3):
            results = []
            for i, n in enumerate(sizes):
                x = rng.uniform(-1, 1, n)
                xt = torch.from_numpy(x).to(tdt).cuda()
                kind = (i + rnd) % 5
                if kind == 3:
                    g.advance(xt)
                    o.advance(x)
                elif kind == 4 and i % 2:
                    y = g.roundtrip(xt)
                    results.append(("y", y, o.roundtrip(x)))
                elif kind == 4:
                    results.append(("rows_host", g.sdft(x), o.sdft(x)))
                else:
                    results.append(("rows", g.sdft(xt), o.sdft(x)))
            g.synchronize()
            for what, got, ref in results:
                got = got.cpu().numpy() if hasattr(got, "cpu") else got
                if what == "y":
                    assert np.abs(got.astype(np.float64) - ref).max() <= (2e-6 if fd == "f64" else 2e-4) * max(np.abs(ref).max(), 1e-3)
                else:
                    assert rel_err(got, ref) <= TOL[fd], (m, rnd, what, rel_err(got, ref))
            if rnd == 0:
                g.set_streaming(2)
            if rnd == 1:
                g.reset()
                o.reset()
        cg, hg, ag, _ = g.state()
        co, ho, ao, _ = o.state()
        assert cg == co and np.array_equal(_bits(hg), _bits(ho)) and rel_err(ag, ao) <= TOL[fd]


@pytest.mark.parametrize("fd", ["f64", "f32"])
def test_streaming_mode_batch_plan_roi_and_side_stream(SDFT, fd):
    """Streaming on a BATCH plan (the hand-over counters count per channel), with a region of interest, queued on
    a torch side stream while the default stream is busy: every channel against its own oracle."""
    import torch
    from oracle import Oracle
    m, ch, hop, calls = 250, 3, 512, 40
    rng = np.random.default_rng(seed_of("stream-batch", fd))
    x = rng.uniform(-1, 1, (ch, hop * calls)).astype(np.float32)
    xt = torch.from_numpy(x).cuda()
    g = SDFT(m, "hamming", 0.5, td="f32", fd=fd, channels=ch)
    g.set_streaming(6)
    g.set_roi(40, 120)
    side = torch.cuda.Stream()
    busy = torch.empty(1 << 26, device="cuda")
    outs = []
    torch.cuda.synchronize()
    with torch.cuda.stream(side):
        for c in range(calls):
            busy.normal_()                                   # unrelated work between the calls, on the same stream
            seg = xt[:, c * hop:(c + 1) * hop].contiguous()
            side.synchronize()                               # the promise: samples complete when the call is issued
            outs.append(g.sdft(seg))
        y = g.isdft(outs[-1])                                # queued behind the calls in the ordinary way
    side.synchronize()
    g._check()
    got = torch.stack(outs, dim=1).cpu().numpy()             # (ch, calls, hop, 120)
    for k in range(ch):
        o = Oracle("f32", fd, m, "hamming", 0.5)
        want = o.sdft(x[k])
        w = want.reshape(calls, hop, m)[:, :, 40:160]
        assert np.abs(got[k] - w).max() <= TOL[fd] * np.abs(want).max(), k
    assert y.shape == (ch, hop)


def test_host_tiling_matches_single_pass(SDFT, monkeypatch):
    """Host destinations are produced in device tiles; tiny tiles must not change a bit."""
    m, n = 64, 5000
    x = np.random.default_rng(41).uniform(-1, 1, n).astype(np.float32)
    g = SDFT(m, "hann", 1, td="f32", fd="f64")
    g.set_chunk(64)
    full = g.sdft(x)
    monkeypatch.setenv("SDFT_B200_TILE_MB", "1")
    g2 = SDFT(m, "hann", 1, td="f32", fd="f64")
    g2.set_chunk(64)
    tiled = g2.sdft(x)
    assert rel_err(tiled, full) <= 1e-13
    y1, y2 = g.isdft(full), g2.isdft(full)
    assert np.array_equal(y1, y2)


def test_pageable_and_pinned_host_buffers_agree(SDFT, monkeypatch):
    """Host-pointer calls: pageable caller memory goes through the library's pinned staging and host copy
    threads, page-locked memory is the DMA target itself; both must deliver the same bytes, for rows out
    (sdft_n) and rows in (isdft_n), across several tiles."""
    m, n = 256, 9000
    monkeypatch.setenv("SDFT_B200_TILE_MB", "4")
    x = np.random.default_rng(43).uniform(-1, 1, n).astype(np.float32)
    a = SDFT(m, "hann", 0.5, td="f32", fd="f64")
    b = SDFT(m, "hann", 0.5, td="f32", fd="f64")
    lib = a._lib
    pageable = a.sdft(x)
    nbytes = n * m * 16
    ptr = lib.sdft_b200_host_alloc(nbytes)
    assert ptr
    try:
        pinned = np.ctypeslib.as_array((ctypes.c_char * nbytes).from_address(ptr)).view(np.complex128).reshape(n, m)
        lib.sdft_b200_f32f64_sdft_n(b._h, n, x.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptr))
        b._check()
        assert np.array_equal(_bits(pageable), _bits(pinned))
        y_pageable = a.isdft(pageable)
        y_pinned = np.empty(n, np.float32)
        lib.sdft_b200_f32f64_isdft_n(b._h, n, ctypes.c_void_p(ptr), y_pinned.ctypes.data_as(ctypes.c_void_p))
        b._check()
        assert np.array_equal(y_pageable, y_pinned)
    finally:
        lib.sdft_b200_host_free(ctypes.c_void_p(ptr))


def test_pageable_staging_under_many_small_calls(SDFT, monkeypatch):
    """Hammer the host copy threads: hundreds of host-pointer calls, each cut into several tiny tiles."""
    m = 64
    monkeypatch.setenv("SDFT_B200_TILE_MB", "1")
    g = SDFT(m, "hann", 1, td="f32", fd="f32")
    r = SDFT(m, "hann", 1, td="f32", fd="f32")
    rng = np.random.default_rng(47)
    import torch
    for k in range(300):
        n = int(rng.integers(1, 6000))
        x = rng.uniform(-1, 1, n).astype(np.float32)
        got = g.sdft(x)
        # same kernels without host staging, cut the way the host path tiles the call (1 MiB of rows)
        rows = (1 << 20) // (m * 8)
        xd = torch.from_numpy(x).cuda()
        want = torch.cat([r.sdft(xd[i:i + rows]) for i in range(0, n, rows)])
        assert np.array_equal(_bits(got), _bits(want.cpu().numpy())), k
        assert np.array_equal(g.isdft(got), r.isdft(want).cpu().numpy()), k


def test_config1_testwav(SDFT, golden_dir):
    """BASELINE config 1: test/test.wav, m=1024, hann, f32 TD / f64 FD, latency 1; whole signal in
    4096-sample calls, against the oracle, the golden rows and the reference's reconstruction SNR."""
    from oracle import Oracle
    g = np.load(os.path.join(golden_dir, "testwav.npz"))
    x = (g["pcm24"].astype(np.float64) / 8388608.0).astype(np.float32)
    n, call, m = int(g["c1_n"]), 4096, 1024
    gpu = SDFT(m, "hann", 1, td="f32", fd="f64")
    cpu = Oracle("f32", "f64", m, "hann", 1.0)
    want_t = list(g["c1_row_t"])
    ys, k, worst = [], 0, 0.0
    for c in range(n // call):
        seg = x[c * call:(c + 1) * call]
        got = gpu.sdft(seg)
        want = cpu.sdft(seg)
        worst = max(worst, rel_err(got, want))
        ys.append(gpu.isdft(got))
        if k < len(want_t) and want_t[k] == c * call + call - 1:
            assert rel_err(got[-1], g["c1_rows"][k]) <= 1e-9
            k += 1
    assert worst <= 1e-9, worst
    y = np.concatenate(ys)
    assert np.abs(y[::int(g["c1_y_stride"])] - g["c1_y_strided"]).max() <= 2e-6
    delay = m - 1
    xd = x[:n - delay].astype(np.float64)
    e = y[delay:].astype(np.float64) - xd
    snr = 10 * np.log10(np.mean(xd ** 2) / np.mean(e ** 2))
    assert abs(snr - float(g["c1_snr_db"])) < 0.01
    # the same signal in ONE call (5.8 GB of rows, device resident): golden rows, samples and SNR again
    import torch
    whole = SDFT(m, "hann", 1, td="f32", fd="f64")
    rows = whole.sdft(torch.from_numpy(x[:n]).cuda())
    for k, t in enumerate(want_t):
        assert rel_err(rows[int(t)].cpu().numpy(), g["c1_rows"][k]) <= 1e-9
    y1 = whole.isdft(rows).cpu().numpy()
    assert np.abs(y1 - y).max() <= 2e-6
    e1 = y1[delay:].astype(np.float64) - xd
    assert abs(10 * np.log10(np.mean(xd ** 2) / np.mean(e1 ** 2)) - float(g["c1_snr_db"])) < 0.01
