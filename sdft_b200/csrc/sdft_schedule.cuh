/*
 * sdft_schedule.cuh -- the phase sources (float table kernel K0, double roots of unity) and the phase at an arbitrary cursor (c/src/sdft/sdft.h:566-576, :584); pulls in the chunk schedule (sdft_chunks.hpp).
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include "sdft_arith.cuh"
#include "sdft_chunks.hpp"

namespace sdftb200
{

/* ------------------------------------------------------------------------------------------------
 * Where a chunk's starting phase P[c][e] = tw[e]^c comes from (the reference's "fiddle" at cursor c,
 * sdft.h:566-576, :584).  Both sources are O(m) .. O(m^2 / stride) bytes, never the O(m^2 / 32) of one row per
 * 32 cursors at every m:
 *
 *   float   must reproduce the reference's SEQUENTIAL float recurrence bit for bit, so the
 *           phase is a table row F0[c / stride][e] (K0 below, built by that recurrence) plus c % stride
 *           rotations.  `stride` is 32 for ordinary sizes and doubles with m so that the table stays within a
 *           fixed budget (sdft_plan.hpp: f0_stride_for); chunk lengths are multiples of the stride, so only the
 *           first chunk of a call ever rotates.
 *   double  needs no bit-exact table (gate 1e-9; an exactly rounded phase is CLOSER to tw^c than the
 *           recurrence's own value): P[c][k] = E[(c k) mod 2m] from ONE table of the 2m roots of unity,
 *           E[j] = exp(-2 pi i j / 2m), shared by all bins -- 32 bytes per bin.  Mirror cells take the
 *           conjugate of their source bin's phase (they are carried with conjugated twiddles).
 * ---------------------------------------------------------------------------------------------- */
template <typename F> struct PhaseSource
{
  const cx<F>* f0;        // float: (rows, cells) table; double: one row, P[0] (1, or 0 for always-zero mirror cells)
  const cx<F>* roots;     // double: E[0 .. 2m); float: unused
  unsigned cells;
  unsigned m;
  unsigned period;        // 2m
  unsigned long long inv_period;   // floor((2^64 - 1) / 2m): (c k) mod 2m by one multiply-high instead of a 64-bit division
  unsigned stride;        // cursors per table row (float)
  int mir_cell[4];        // mirror cells (sdft.h:589-595): cell index, source CELL (-1: always zero), conjugated?
  int mir_src[4];
  int mir_conj[4];
};

/* K0  float phase table: F0[row][e] = P[row * stride][e] by the sequential recurrence (sdft.h:584) */
template <typename F>
__global__ void phase_table_kernel(const cx<F>* __restrict__ tw_ext, const cx<F>* __restrict__ p0_ext,
                                   cx<F>* __restrict__ f0, unsigned cells, unsigned period, unsigned stride)
{
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cells) return;
  const cx<F> w = tw_ext[e];
  cx<F> p = p0_ext[e];
  for (unsigned c = 0; c < period; ++c)
  {
    if (c % stride == 0) f0[(size_t)(c / stride) * cells + e] = p;
    p = Arith<F>::rotate(p, w);
  }
}

/* phase of cell e at cursor c */
template <typename F>
__device__ __forceinline__ cx<F> phase_at(const PhaseSource<F>& s, int e, unsigned cursor, cx<F> w)
{
  if constexpr (sizeof(F) == sizeof(double))
  {
    (void)w;
    int cell = e;
    bool conj = false;
    if (e < 2 || e >= (int)s.m + 2)
    {
      cell = -1;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (s.mir_cell[q] == e) { cell = s.mir_src[q]; conj = s.mir_conj[q] != 0; }
      if (cell < 0)
      {
        cx<F> z; z.r = (F)0; z.i = (F)0;
        return z;
      }
    }
    const unsigned long long x = (unsigned long long)cursor * (unsigned long long)(cell - 2);    // < 2^61
    unsigned long long j = x - __umul64hi(x, s.inv_period) * s.period;                            // x mod 2m, up to + 2m
    if (j >= s.period) j -= s.period;
    if (j >= s.period) j -= s.period;
    cx<F> p = s.roots[j];
    if (conj) p.i = -p.i;
    return p;
  }
  else
  {
    /* the same values the sequential recurrence of the reference produces, bit for bit */
    cx<F> p = s.f0[(size_t)(cursor / s.stride) * s.cells + e];
    const unsigned steps = cursor % s.stride;
    for (unsigned i = 0; i < steps; ++i) p = Arith<F>::rotate(p, w);
    return p;
  }
}

/* phase right after the period's restart: P[0] = 1 (0 for an always-zero mirror cell) */
template <typename F>
__device__ __forceinline__ cx<F> phase_restart(const PhaseSource<F>& s, int e)
{
  return s.f0[e];
}

/* introspection (sdft_b200_get_state): the modulation phase of every bin at `cursor` */
template <typename F>
__global__ void phase_at_kernel(const cx<F>* __restrict__ tw_ext, const PhaseSource<F> src, cx<F>* __restrict__ out, unsigned cursor)
{
  const unsigned e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= src.cells) return;
  out[e] = phase_at<F>(src, (int)e, cursor, tw_ext[e]);
}


}  // namespace sdftb200
