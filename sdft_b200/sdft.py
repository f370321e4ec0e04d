"""
Python ``SDFT`` class over the C-ABI CUDA library.

Mirrors the reference's NumPy class (python/src/sdft/sdft.py:25-203): same constructor arguments, same
``size`` / ``window`` / ``latency`` attributes, ``reset()``, ``sdft(samples) -> (samples, bins)`` and
``isdft(dfts) -> (samples,)``.  The arithmetic is NOT done here: every call goes through
``libsdft_b200.so`` (include/sdft_b200.h) and runs on the GPU.  There is no CPU fallback.

Differences from the reference class, all additive:
  * ``td`` / ``fd`` select the time/frequency-domain precision ("f32" or "f64"); the default
    (f64, f64) is what the reference's float64/complex128 NumPy code computes in.
  * CUDA tensors (torch) are accepted and returned without leaving the device.
  * the plan follows the C implementation exactly, including the periodic modulation restart that
    makes it endlessly stable (c/src/sdft/sdft.h:566-576); the reference Python class lets its phase
    offset grow without bound (sdft.py:101).
"""
import ctypes

import numpy as np

from . import _lib

WINDOWS = ("boxcar", "hann", "hamming", "blackman")
_NP_TD = {"f32": np.float32, "f64": np.float64}
_NP_FD = {"f32": np.complex64, "f64": np.complex128}


def _window_id(window):
    if isinstance(window, str):
        w = window.lower()
        # the reference matches by substring (sdft.py:160-186) and falls back to boxcar
        for i in (1, 2, 3):
            if w and w in WINDOWS[i]:
                return i
        return 0
    return int(window)


def _is_torch_cuda(x):
    return type(x).__module__.startswith("torch") and hasattr(x, "is_cuda") and x.is_cuda


def _device_view(x):
    """CUDA tensors pass through; any other object exposing __cuda_array_interface__ (CuPy, Numba, ...) is
    wrapped zero-copy as a torch tensor; host data is returned unchanged."""
    if _is_torch_cuda(x) or not hasattr(x, "__cuda_array_interface__"):
        return x
    import torch
    return torch.as_tensor(x, device="cuda")


class SDFT:
    """Sliding Discrete Fourier Transform (SDFT) on a B200."""

    def __init__(self, dftsize, window="hann", latency=1, td="f64", fd="f64", channels=1):
        self.size = int(dftsize)
        self.window = window
        self.latency = latency
        self.td, self.fd = td, fd
        self.channels = int(channels)
        self._rowlen = self.size        # bins per row: the region of interest, see set_roi
        self._sfx = td + fd
        self._lib = _lib.load()
        self._f = lambda name: _lib.fn(self._lib, self._sfx, name)
        self._h = self._f("alloc_batch")(self.size, _window_id(window), float(latency), self.channels)
        if not self._h:
            msg = self._lib.sdft_b200_last_error_string(None)
            raise RuntimeError("sdft_b200: plan allocation failed: %s" % (msg.decode() if msg else "?"))
        self._h = ctypes.c_void_p(self._h)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                self._f("free")(h)
            except Exception:
                pass

    # ---- helpers -------------------------------------------------------------------------------
    def _check(self):
        code = self._lib.sdft_b200_last_error(self._h)
        if code:
            raise RuntimeError("sdft_b200: %s" % self._lib.sdft_b200_last_error_string(self._h).decode())

    def _use_torch_stream(self, like=None):
        """Queue on torch's current stream OF THE PLAN'S DEVICE; `like` (a CUDA tensor handed to the call)
        must live on that device -- the kernels would otherwise dereference another GPU's memory."""
        import torch
        dev = int(self._lib.sdft_b200_device(self._h))
        if like is not None:
            assert like.device.index == dev, f'Expected a tensor on cuda:{dev} (the plan\'s device), got {like.device}!'
        self._lib.sdft_b200_set_stream(self._h, ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream))

    def _check_batch(self, shape, trailing):
        """(channels, samples[, bins]) on batch plans, (samples[, bins]) otherwise."""
        if self.channels == 1:
            assert len(shape) == trailing or (len(shape) == trailing + 1 and shape[0] == 1), \
                f'Expected {trailing}D input for a single-channel plan, got {tuple(shape)}!'
        else:
            assert len(shape) == trailing + 1 and shape[0] == self.channels, \
                f'Expected ({self.channels}, ...) for a {self.channels}-channel plan, got {tuple(shape)}!'

    def synchronize(self):
        self._lib.sdft_b200_synchronize(self._h)
        self._check()

    @property
    def launches(self):
        return int(self._lib.sdft_b200_launch_count(self._h))

    # ---- reference API ---------------------------------------------------------------------------
    def reset(self):
        """Reset this SDFT plan to its initial state (sdft.py:67-74)."""
        self._f("reset")(self._h)
        self._check()

    def sdft(self, samples, out=None):
        """Estimate the DFT matrix for the given sample array (sdft.py:76-120).

        Returns (samples, bins); for a batch plan the input is (channels, samples) and the result
        (channels, samples, bins).  `out` (CUDA tensors only) reuses a caller-owned result tensor."""
        samples = _device_view(samples)
        if _is_torch_cuda(samples):
            import torch
            x = samples.to(torch.float32 if self.td == "f32" else torch.float64).contiguous()
            self._check_batch(x.shape, 1)
            n = x.shape[-1]
            shape = (n, self._rowlen) if self.channels == 1 else (self.channels, n, self._rowlen)
            fdt = torch.complex64 if self.fd == "f32" else torch.complex128
            if out is None:
                out = torch.empty(shape, dtype=fdt, device=x.device)
            else:
                assert out.is_cuda and out.is_contiguous() and out.dtype == fdt and out.numel() >= n * self._rowlen * self.channels
                assert out.device == x.device
                out = out.view(-1)[: n * self._rowlen * self.channels].view(shape)
            self._use_torch_stream(x)
            self._f("sdft_batch")(self._h, n, ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()))
            self._check()
            return out
        x = np.ascontiguousarray(np.atleast_1d(samples), dtype=_NP_TD[self.td])
        if self.channels == 1:
            assert x.ndim == 1, f'Expected 1D array (samples,), got {x.shape}!'
        else:
            assert x.ndim == 2 and x.shape[0] == self.channels, f'Expected (channels,samples), got {x.shape}!'
        n = x.shape[-1]
        shape = (n, self._rowlen) if self.channels == 1 else (self.channels, n, self._rowlen)
        out = np.empty(shape, _NP_FD[self.fd])
        self._f("sdft_batch")(self._h, n, x.ctypes.data_as(ctypes.c_void_p), out.ctypes.data_as(ctypes.c_void_p))
        self._check()
        return out

    def isdft(self, dfts):
        """Synthesize the sample array from the given DFT matrix (sdft.py:122-145)."""
        dfts = _device_view(dfts)
        if _is_torch_cuda(dfts):
            import torch
            d = dfts.to(torch.complex64 if self.fd == "f32" else torch.complex128).contiguous()
            assert d.shape[-1] == self._rowlen, f'Expected (samples,frequencies), got {tuple(d.shape)}!'
            self._check_batch(d.shape, 2)
            n = d.shape[-2]
            shape = (n,) if self.channels == 1 else (self.channels, n)
            y = torch.empty(shape, dtype=torch.float32 if self.td == "f32" else torch.float64, device=d.device)
            self._use_torch_stream(d)
            self._f("isdft_batch")(self._h, n, ctypes.c_void_p(d.data_ptr()), ctypes.c_void_p(y.data_ptr()))
            self._check()
            return y
        d = np.ascontiguousarray(np.atleast_2d(dfts), dtype=_NP_FD[self.fd])
        assert d.shape[-1] == self._rowlen, f'Expected (samples,frequencies), got {d.shape}!'
        n = d.shape[-2]
        shape = (n,) if self.channels == 1 else (self.channels, n)
        y = np.empty(shape, _NP_TD[self.td])
        self._f("isdft_batch")(self._h, n, d.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p))
        self._check()
        return y

    # ---- extensions ------------------------------------------------------------------------------
    def advance(self, samples):
        """Update the analysis state with ``samples`` without producing rows (time-shard priming)."""
        samples = _device_view(samples)
        if _is_torch_cuda(samples):
            import torch
            x = samples.to(torch.float32 if self.td == "f32" else torch.float64).contiguous()
            self._check_batch(x.shape, 1)
            self._use_torch_stream(x)
            self._f("advance")(self._h, x.shape[-1], ctypes.c_void_p(x.data_ptr()))
        else:
            x = np.ascontiguousarray(np.atleast_1d(samples), dtype=_NP_TD[self.td])
            self._f("advance")(self._h, x.shape[-1], x.ctypes.data_as(ctypes.c_void_p))
        self._check()

    def roundtrip(self, samples, gains=None):
        """isdft(sdft(samples)) in one fused kernel: the DFT matrix never exists in memory.
        `gains` (bins,) complex: every DFT row is multiplied with it before synthesis (a static
        spectral filter between analysis and synthesis)."""
        gp = None
        if gains is not None:
            gv = np.ascontiguousarray(gains, dtype=_NP_FD[self.fd])
            assert gv.shape == (self.size,), f'Expected (frequencies,), got {gv.shape}!'
            gp = gv.ctypes.data_as(ctypes.c_void_p)

        def call(n, xp, yp):
            if gp is None:
                self._f("roundtrip_n")(self._h, n, xp, yp)
            else:
                self._f("roundtrip_gain_n")(self._h, n, xp, yp, gp)
            self._check()

        samples = _device_view(samples)
        if _is_torch_cuda(samples):
            import torch
            x = samples.to(torch.float32 if self.td == "f32" else torch.float64).contiguous()
            self._check_batch(x.shape, 1)
            y = torch.empty_like(x)
            self._use_torch_stream(x)
            call(x.shape[-1], ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(y.data_ptr()))
            return y
        x = np.ascontiguousarray(np.atleast_1d(samples), dtype=_NP_TD[self.td])
        y = np.empty_like(x)
        call(x.shape[-1], x.ctypes.data_as(ctypes.c_void_p), y.ctypes.data_as(ctypes.c_void_p))
        return y

    def convolve(self, x):
        """Window the specified DFT matrix (sdft.py:146-203): window(x) / bins, on the GPU."""
        x = _device_view(x)
        if _is_torch_cuda(x):
            import torch
            d = torch.atleast_2d(x).to(torch.complex64 if self.fd == "f32" else torch.complex128).contiguous()
            assert d.shape[-1] == self.size, f'Expected (samples,frequencies), got {tuple(d.shape)}!'
            out = torch.empty_like(d)
            self._use_torch_stream(d)
            self._f("convolve_n")(self._h, d.numel() // self.size // self.channels, ctypes.c_void_p(d.data_ptr()),
                                  ctypes.c_void_p(out.data_ptr()))
            self._check()
            return out
        d = np.ascontiguousarray(np.atleast_2d(x), dtype=_NP_FD[self.fd])
        assert d.shape[-1] == self.size, f'Expected (samples,frequencies), got {d.shape}!'
        out = np.empty_like(d)
        self._f("convolve_n")(self._h, d.size // self.size // self.channels, d.ctypes.data_as(ctypes.c_void_p),
                              out.ctypes.data_as(ctypes.c_void_p))
        self._check()
        return out

    def twiddles(self):
        a = np.empty(self.size, _NP_FD[self.fd])
        s = np.empty(self.size, _NP_FD[self.fd])
        self._lib.sdft_b200_get_twiddles(self._h, a.ctypes.data_as(ctypes.c_void_p), s.ctypes.data_as(ctypes.c_void_p))
        self._check()
        return a, s

    def state(self, channel=0):
        cur = ctypes.c_size_t(0)
        hist = np.empty(2 * self.size, _NP_TD[self.td])
        acc = np.empty(self.size, _NP_FD[self.fd])
        ph = np.empty(self.size, _NP_FD[self.fd])
        self._lib.sdft_b200_get_state(self._h, channel, ctypes.byref(cur), hist.ctypes.data_as(ctypes.c_void_p),
                                      acc.ctypes.data_as(ctypes.c_void_p), ph.ctypes.data_as(ctypes.c_void_p))
        self._check()
        return int(cur.value), hist, acc, ph

    def set_state(self, cursor, history=None, accumulators=None, channel=0):
        """Import plan state (the counterpart of ``state()``): cursor, the last 2*size samples (oldest first)
        and/or the per-bin accumulators; used to start a time shard exactly where a continuous run would be."""
        hp = ap = None
        if history is not None:
            h = np.ascontiguousarray(history, dtype=_NP_TD[self.td])
            assert h.shape == (2 * self.size,)
            hp = h.ctypes.data_as(ctypes.c_void_p)
        if accumulators is not None:
            a = np.ascontiguousarray(accumulators, dtype=_NP_FD[self.fd])
            assert a.shape == (self.size,)
            ap = a.ctypes.data_as(ctypes.c_void_p)
        self._lib.sdft_b200_set_state(self._h, channel, int(cursor), hp, ap)
        self._check()

    def set_roi(self, first=0, count=0):
        """Region of interest: from now on ``sdft`` returns and ``isdft`` takes rows of the bins
        [first, first + count) only; count=0 restores the whole spectrum.  Bins outside still take part in the
        state update, they just cost no row bandwidth."""
        if self._lib.sdft_b200_set_roi(self._h, int(first), int(count)):
            raise ValueError("sdft_b200: region of interest outside [0, %d)" % self.size)
        self._rowlen = int(count) if 0 < int(count) < self.size else self.size

    def set_streaming(self, depth):
        """Streaming mode for endless sequences of short calls on device buffers (the hop loop of
        test/test.c:69-83): up to `depth` consecutive calls may be in flight at once.  See
        sdft_b200_set_streaming in include/sdft_b200.h for what the caller promises; depth <= 1 restores plain
        stream order."""
        self._lib.sdft_b200_set_streaming(self._h, int(depth))
        self._check()

    def streaming(self, depth=8):
        """``with plan.streaming(8): ...`` -- streaming mode for the block, plain stream order again afterwards
        (leaving the block waits for the calls queued inside it)."""
        import contextlib

        @contextlib.contextmanager
        def scope():
            self.set_streaming(depth)
            try:
                yield self
            finally:
                self.set_streaming(1)
        return scope()

    def sdft_hops(self, samples, hop, out):
        """The reference drivers' hop loop issued from inside the library: `out[h] = sdft(samples[h*hop:(h+1)*hop])`
        for CUDA tensors `samples` (nhops*hop,) and `out` (nhops, hop, bins)."""
        assert _is_torch_cuda(samples) and _is_torch_cuda(out) and self.channels == 1
        assert samples.is_contiguous() and out.is_contiguous() and samples.numel() % hop == 0
        nhops = samples.numel() // hop
        assert out.numel() == nhops * hop * self._rowlen
        self._use_torch_stream(samples)
        self._f("sdft_hops")(self._h, nhops, int(hop), ctypes.c_void_p(samples.data_ptr()), ctypes.c_void_p(out.data_ptr()),
                             int(hop) * self._rowlen)
        self._check()
        return out

    def set_chunk(self, chunk):
        self._lib.sdft_b200_set_chunk(self._h, int(chunk))
