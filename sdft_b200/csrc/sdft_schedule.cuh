/*
 * sdft_schedule.cuh -- the phase table: kernel K0 and the phase at an arbitrary cursor (c/src/sdft/sdft.h:566-576, :584); pulls in the chunk schedule (sdft_chunks.hpp).
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include "sdft_arith.cuh"
#include "sdft_chunks.hpp"

namespace sdftb200
{

/* ------------------------------------------------------------------------------------------------
 * K0  phase table: F0[row][e] = P[row * kF0Stride][e] by the sequential recurrence (sdft.h:584)
 * ---------------------------------------------------------------------------------------------- */
template <typename F>
__global__ void phase_table_kernel(const cx<F>* __restrict__ tw_ext, const cx<F>* __restrict__ p0_ext,
                                   cx<F>* __restrict__ f0, unsigned cells, unsigned period)
{
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cells) return;
  const cx<F> w = tw_ext[e];
  cx<F> p = p0_ext[e];
  for (unsigned c = 0; c < period; ++c)
  {
    if (c % kF0Stride == 0) f0[(size_t)(c / kF0Stride) * cells + e] = p;
    p = Arith<F>::rotate(p, w);
  }
}

/* phase at an arbitrary cursor: table row + (cursor % 32) rotations -- the same values the sequential
 * recurrence of the reference produces (bit-identical for float) */
template <typename F>
__device__ __forceinline__ cx<F> phase_at(const cx<F>* __restrict__ f0, unsigned cells, int e, unsigned cursor, cx<F> w)
{
  cx<F> p = f0[(size_t)(cursor / kF0Stride) * cells + e];
  const unsigned steps = cursor % kF0Stride;
  for (unsigned i = 0; i < steps; ++i) p = Arith<F>::rotate(p, w);
  return p;
}

/* introspection (sdft_b200_get_state): the modulation phase of every bin at `cursor` */
template <typename F>
__global__ void phase_at_kernel(const cx<F>* __restrict__ tw_ext, const cx<F>* __restrict__ f0, cx<F>* __restrict__ out,
                                unsigned cells, unsigned cursor)
{
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cells) return;
  out[e] = phase_at<F>(f0, cells, (int)e, cursor, tw_ext[e]);
}


}  // namespace sdftb200
