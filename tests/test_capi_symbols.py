"""CPU-side checks of the boundary: the library builds, loads and exports every declared symbol."""
import ctypes
import os
import re

from sdft_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_builds_and_exports_every_symbol():
    build.build()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in _lib.exported_symbols():
        assert hasattr(lib, name), name


def test_header_and_binding_agree():
    """Every SDFT_B200_API function in include/sdft_b200.h is bound by sdft_b200/_lib.py."""
    text = open(os.path.join(ROOT, "include", "sdft_b200.h")).read()
    typed = set(re.findall(r"sdft_b200_##SFX##_(\w+)\(", text))
    assert typed == set(_lib.TYPED), typed ^ set(_lib.TYPED)
    untyped = set(re.findall(r"SDFT_B200_API [\w \*]+?\b(sdft_b200_\w+)\(", text))
    untyped = {u for u in untyped if "##" not in u}
    assert untyped == set(_lib.UNTYPED), untyped ^ set(_lib.UNTYPED)


def test_no_cpu_fallback_without_device():
    """Without a CUDA device plan allocation fails loudly instead of computing on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    lib = _lib.load()
    h = lib.sdft_b200_f32f64_alloc(64)
    assert not h
    assert lib.sdft_b200_last_error(None) != 0
    assert b"no CUDA device" in lib.sdft_b200_last_error_string(None) or lib.sdft_b200_last_error(None) != 0


def test_product_does_not_touch_the_oracle():
    """The oracle is test infrastructure: nothing under sdft_b200/ or include/ may reference it."""
    for base in ("sdft_b200", "include"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, base)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp")):
                    src = open(os.path.join(dirpath, f), errors="ignore").read()
                    assert "oracle" not in src.lower() or f == "_never_", (dirpath, f)
