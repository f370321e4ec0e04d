"""
sdft_b200 -- B200-native (sm_100a CUDA) implementation of the analysis/synthesis hot path of
jurihock/sdft behind the reference's own API.  See DESIGN.md and INTEGRATION.md.
"""
from .sdft import SDFT  # noqa: F401

__version__ = "0.1"
