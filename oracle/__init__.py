"""
oracle/ -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

ctypes loaders for
  * ``Oracle``  -- the CPU restatement built from oracle/sdft_oracle.c (``libsdft_oracle.so``);
  * ``Ref``     -- the UNMODIFIED reference header (c/src/sdft/sdft.h) compiled by oracle/Makefile into
                   oracle/_ref/libsdft_ref_<td><fd>[_fast].so.

Only tests/, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may
import this package; the product (``sdft_b200``) never does.

Both classes share one interface: ``sdft(x) -> (n, m) complex``, ``isdft(dfts) -> (n,) real``,
``reset()``, ``twiddles() -> (analysis, synthesis)``, ``state() -> (cursor, history, acc, phase)``.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
REF_ROOT = "/root/reference"

WINDOWS = {"boxcar": 0, "hann": 1, "hamming": 2, "blackman": 3}
_NP = {"f32": np.float32, "f64": np.float64}
_NPX = {"f32": np.complex64, "f64": np.complex128}


def window_id(window):
    return WINDOWS[window] if isinstance(window, str) else int(window)


def build(want_ref=True):
    """Compile the restatement (always) and, when the reference tree is present, oracle/_ref."""
    targets = ["oracle"]
    if want_ref and os.path.exists(os.path.join(REF_ROOT, "c/src/sdft/sdft.h")):
        targets.append("ref")
    subprocess.run(["make", "-s", "-C", HERE] + targets, check=True)


def _cpu_has_avx512():
    """x86-64-v4: what the _fast512 timing build of the reference needs"""
    try:
        flags = next(l for l in open("/proc/cpuinfo") if l.startswith("flags")).split()
    except (OSError, StopIteration):
        return False
    return all(f in flags for f in ("avx512f", "avx512bw", "avx512cd", "avx512dq", "avx512vl"))


def fast_suffix():
    """Timing build of the reference for THIS host: x86-64-v4 when the CPU has AVX-512 and the library was built,
    else x86-64-v3 (what -O3 -march=native, BASELINE.md, comes to on the two kinds of host this runs on)."""
    if _cpu_has_avx512() and os.path.exists(os.path.join(REF_DIR, "libsdft_ref_f32f64_fast512.so")):
        return "_fast512"
    return "_fast"


def fast_flags():
    return "-O3 -march=x86-64-v4" if fast_suffix() == "_fast512" else "-O3 -march=x86-64-v3"


def have_ref(fast=False):
    name = "libsdft_ref_f32f64%s.so" % ("_fast" if fast else "")
    return os.path.exists(os.path.join(REF_DIR, name))


_libs = {}


def _load(path):
    if path not in _libs:
        _libs[path] = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _libs[path]


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _Base:
    td = fd = None

    def _setup(self, td, fd, m, window, latency):
        self.td, self.fd = td, fd
        self.m = int(m)
        self.window = window_id(window)
        self.latency = float(latency)
        self.td_np, self.fd_np, self.fdx_np = _NP[td], _NP[fd], _NPX[fd]

    def _as_td(self, x):
        return np.ascontiguousarray(np.atleast_1d(x), dtype=self.td_np)

    def _as_fdx(self, d):
        d = np.ascontiguousarray(np.atleast_2d(d), dtype=self.fdx_np)
        assert d.shape[1] == self.m
        return d


class Oracle(_Base):
    """CPU restatement (oracle/sdft_oracle_impl.h)."""

    def __init__(self, td, fd, m, window="hann", latency=1.0):
        path = os.path.join(HERE, "libsdft_oracle.so")
        if not os.path.exists(path):
            build(want_ref=False)
        self.lib = _load(path)
        self._setup(td, fd, m, window, latency)
        self.sfx = "oracle_%s%s_" % (td, fd)
        f = self._fn("alloc", ctypes.c_void_p, [ctypes.c_size_t, ctypes.c_int, ctypes.c_double])
        self.h = ctypes.c_void_p(f(self.m, self.window, self.latency))

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, self.sfx + name)
        f.restype, f.argtypes = restype, argtypes
        return f

    def __del__(self):
        try:
            self._fn("free", None, [ctypes.c_void_p])(self.h)
        except Exception:
            pass

    def reset(self):
        self._fn("reset", None, [ctypes.c_void_p])(self.h)

    def sdft(self, x):
        x = self._as_td(x)
        out = np.empty((x.size, self.m), self.fdx_np)
        self._fn("sdft_n", None, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p])(
            self.h, x.size, _ptr(x), _ptr(out))
        return out

    def isdft(self, dfts):
        d = self._as_fdx(dfts)
        y = np.empty(d.shape[0], self.td_np)
        self._fn("isdft_n", None, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p])(
            self.h, d.shape[0], _ptr(d), _ptr(y))
        return y

    def advance(self, x):
        """State update only (same arithmetic as sdft, no rows): used to walk long signals."""
        x = self._as_td(x)
        self._fn("advance_n", None, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p])(self.h, x.size, _ptr(x))

    def clone(self, window=None):
        """Deep copy of the plan and its state, optionally with another window."""
        other = Oracle.__new__(Oracle)
        other.lib = self.lib
        other._setup(self.td, self.fd, self.m, self.window if window is None else window, self.latency)
        other.sfx = self.sfx
        f = self._fn("clone", ctypes.c_void_p, [ctypes.c_void_p, ctypes.c_int])
        other.h = ctypes.c_void_p(f(self.h, other.window))
        return other

    def roundtrip(self, x):
        """isdft(sdft(x)) sample by sample without keeping the matrix."""
        x = self._as_td(x)
        y = np.empty(x.size, self.td_np)
        row = np.empty(2 * self.m, self.fd_np)
        self._fn("roundtrip_n", None, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p])(
            self.h, x.size, _ptr(x), _ptr(y), _ptr(row))
        return y

    def twiddles(self):
        a = np.empty(self.m, self.fdx_np)
        s = np.empty(self.m, self.fdx_np)
        self._fn("get_twiddles", None, [ctypes.c_void_p] * 3)(self.h, _ptr(a), _ptr(s))
        return a, s

    def state(self):
        hist = np.empty(2 * self.m, self.td_np)
        acc = np.empty(self.m, self.fdx_np)
        ph = np.empty(self.m, self.fdx_np)
        self._fn("get_state", None, [ctypes.c_void_p] * 4)(self.h, _ptr(hist), _ptr(acc), _ptr(ph))
        cur = self._fn("cursor", ctypes.c_size_t, [ctypes.c_void_p])(self.h)
        return int(cur), hist, acc, ph


class Ref(_Base):
    """The reference's own C implementation (c/src/sdft/sdft.h via oracle/ref_shim.c)."""

    def __init__(self, td, fd, m, window="hann", latency=1.0, fast=False):
        path = os.path.join(REF_DIR, "libsdft_ref_%s%s%s.so" % (td, fd, fast_suffix() if fast else ""))
        if not os.path.exists(path):
            build(want_ref=True)
        if not os.path.exists(path):
            raise FileNotFoundError(path)
        self.lib = _load(path)
        self._setup(td, fd, m, window, latency)
        f = self._fn("ref_alloc", ctypes.c_void_p, [ctypes.c_size_t, ctypes.c_int, ctypes.c_double])
        self.h = ctypes.c_void_p(f(self.m, self.window, self.latency))
        assert self._fn("ref_td_size", ctypes.c_size_t, [])() == np.dtype(self.td_np).itemsize
        assert self._fn("ref_fd_size", ctypes.c_size_t, [])() == np.dtype(self.fd_np).itemsize

    def _fn(self, name, restype, argtypes):
        f = getattr(self.lib, name)
        f.restype, f.argtypes = restype, argtypes
        return f

    def __del__(self):
        try:
            self._fn("ref_free", None, [ctypes.c_void_p])(self.h)
        except Exception:
            pass

    def reset(self):
        self._fn("ref_reset", None, [ctypes.c_void_p])(self.h)

    def sdft(self, x):
        x = self._as_td(x)
        out = np.empty((x.size, self.m), self.fdx_np)
        self._fn("ref_sdft_n", None, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p])(
            self.h, x.size, _ptr(x), _ptr(out))
        return out

    def isdft(self, dfts):
        d = self._as_fdx(dfts)
        y = np.empty(d.shape[0], self.td_np)
        self._fn("ref_isdft_n", None, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p])(
            self.h, d.shape[0], _ptr(d), _ptr(y))
        return y

    def twiddles(self):
        a = np.empty(self.m, self.fdx_np)
        s = np.empty(self.m, self.fdx_np)
        self._fn("ref_get_twiddles", None, [ctypes.c_void_p] * 3)(self.h, _ptr(a), _ptr(s))
        return a, s

    def state(self):
        hist = np.empty(2 * self.m, self.td_np)
        acc = np.empty(self.m, self.fdx_np)
        ph = np.empty(self.m, self.fdx_np)
        cur = self._fn("ref_get_state", ctypes.c_size_t, [ctypes.c_void_p] * 4)(
            self.h, _ptr(hist), _ptr(acc), _ptr(ph))
        return int(cur), hist, acc, ph


def cpu_reference(td, fd, m, window="hann", latency=1.0, fast=False):
    """Best available CPU implementation: the compiled reference if present, else the restatement.
    Returns (object, kind) with kind in {"reference", "port"}."""
    if have_ref(fast):
        return Ref(td, fd, m, window, latency, fast=fast), "reference"
    return Oracle(td, fd, m, window, latency), "port"
