/*
 * sdft_peak.cuh -- measurement kernels for the roofline denominators ("plus a same-box pure-store
 * kernel"): a pure streaming STORE with exactly the row kernel's store instruction and cache policy
 * (sdft_lane.cuh: store_group, 32 bytes per lane, consecutive lanes on consecutive groups), a pure streaming READ
 * with the synthesis kernel's load (sdft_synth.cuh: load_stream), and their combination, a copy.  They compute
 * nothing of the transform; bench.py times them in the same run as the row kernel so that `roofline.frac` can be
 * quoted against a write ceiling and a read ceiling of the same board at the same moment, next to the read+write
 * copy peak of MEASURED_PEAKS.json.  Also a DFMA loop for the FP64 issue ceiling of the fused round trip.
 * Part of the sm_100a kernels of libsdft_b200.so.
 */
#pragma once

#include "sdft_lane.cuh"
#include "sdft_synth.cuh"

namespace sdftb200
{

enum { PEAK_STORE = 0, PEAK_READ = 1, PEAK_COPY = 2 };

/* `groups32` 32-byte groups; grid-stride, a warp covers 1 KiB per iteration like one row segment of the row
 * kernel's wide double geometry covers 2 KiB */
template <int KIND>
__global__ void __launch_bounds__(256) peak_stream_kernel(double* __restrict__ dst, const double* __restrict__ src,
                                                          unsigned long long groups32, double* __restrict__ sink)
{
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  double acc = 0.0;
  cx<double> v[2];
  v[0].r = (double)threadIdx.x; v[0].i = 1.0; v[1].r = 2.0; v[1].i = (double)blockIdx.x;
  for (; g < groups32; g += stride)
  {
    if (KIND != PEAK_STORE)
    {
      const cx<double>* s = reinterpret_cast<const cx<double>*>(src) + 2 * g;
      v[0] = load_stream<double>(s);
      v[1] = load_stream<double>(s + 1);
      if (KIND == PEAK_READ) acc += (v[0].r + v[0].i) + (v[1].r + v[1].i);
    }
    if (KIND != PEAK_READ) store_group(reinterpret_cast<cx<double>*>(dst) + 2 * g, v);
  }
  if (KIND == PEAK_READ && acc == 123.456) *sink = acc;      // keeps the loads alive
}

/* FP64 issue ceiling: `iters` x 8 independent DFMA chains per thread, nothing else in the loop */
__global__ void __launch_bounds__(256) peak_dfma_kernel(double* __restrict__ sink, unsigned iters, double a, double b)
{
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 4
  for (unsigned i = 0; i < iters; ++i)
  {
    x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
    x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) *sink = s;
}

}  // namespace sdftb200
