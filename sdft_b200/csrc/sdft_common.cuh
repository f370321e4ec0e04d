/*
 * sdft_common.cuh -- constants, the complex value type and the warp geometries shared by all kernels.
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sdftb200
{


constexpr int kF0Stride = 32;     // finest grid of the float phase table (P at every 32nd cursor) and granularity of chunk lengths
constexpr int kMaxChunk = 1024;   // longest chunk the kernels accept (samples)
constexpr int kAutoChunk = 512;   // longest chunk the heuristic picks (measured best on B200)

template <typename F> struct cx { F r, i; };

/* Work geometry of the emit warps.  A lane owns CPL consecutive cells and stores them as 32-byte
 * groups of GROUP cells; the halo on either side of a warp is one group wide (>= the 2 cells the
 * Blackman taps need), which keeps every group store 32-byte aligned. */
enum { GEO_WIDE = 0, GEO_NARROW = 1 };
template <typename F, int GEO> struct Geo;
template <> struct Geo<double, GEO_WIDE>   { enum { CPL = 4, GROUP = 2, WC = 32 * 4 }; };
template <> struct Geo<float, GEO_WIDE>    { enum { CPL = 8, GROUP = 4, WC = 32 * 8 }; };
/* narrow warps (one 32-byte store group per lane) for short calls: twice the warps, half the work per
 * time step each -- a short call is bound by the latency of its L sequential steps, not by bandwidth */
template <> struct Geo<double, GEO_NARROW> { enum { CPL = 2, GROUP = 2, WC = 32 * 2 }; };
template <> struct Geo<float, GEO_NARROW>  { enum { CPL = 4, GROUP = 4, WC = 32 * 4 }; };


}  // namespace sdftb200
