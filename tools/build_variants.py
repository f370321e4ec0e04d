"""Builds experimental variants of libsdft_b200.so next to the production one (git-ignored *.so):
    python tools/build_variants.py mb5:-DSDFT_B200_MINBLOCKS=5 h1:-DSDFT_B200_HORNER1
Each variant lands in sdft_b200/libsdft_b200_<tag>.so; run with SDFT_B200_LIB=<that path>."""
import os
import sys
from concurrent.futures import ThreadPoolExecutor

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sdft_b200 import build as b


def one(spec):
    tag, _, flags = spec.partition(":")
    out = os.path.join(b.HERE, "libsdft_b200_%s.so" % tag)
    b.build(out=out, extra=[f for f in flags.split(",") if f])
    return out


if __name__ == "__main__":
    with ThreadPoolExecutor(4) as ex:
        for path in ex.map(one, sys.argv[1:]):
            print(path)
