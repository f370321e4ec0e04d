/*
 * sdft_schedule.cuh -- the chunk schedule of a call (identical host/device code) and the phase table kernel K0 (c/src/sdft/sdft.h:566-576, :584).
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include "sdft_arith.cuh"

namespace sdftb200
{

/* ------------------------------------------------------------------------------------------------
 * chunk schedule: identical on host and device
 * ---------------------------------------------------------------------------------------------- */
struct Schedule
{
  unsigned long long cursor;   // cursor before the first sample of the call, 0..2m-1
  unsigned long long n;        // samples in the call
  unsigned period;             // 2m
  unsigned chunk;              // L, multiple of kF0Stride
  unsigned per_period;         // ceil(period / L)
  unsigned first_slot;         // cursor / L
  unsigned nchunks;            // chunks in this call
};

struct ChunkSpan
{
  unsigned long long t0;  // first sample (index inside the call)
  unsigned len;           // samples in the chunk (1..L)
  unsigned cursor0;       // cursor before the chunk's first sample: row cursor0/32 of the phase table
                          // plus cursor0%32 rotations give its phase (0 rotations except for chunk 0)
  bool first;             // chunk 0 of the call
  bool wraps;             // last step is the period's last step (cursor 2m-1): phase restarts
};

__host__ __device__ inline Schedule make_schedule(unsigned long long cursor, unsigned long long n,
                                                  unsigned m, unsigned chunk)
{
  Schedule s;
  s.cursor = cursor;
  s.n = n;
  s.period = 2u * m;
  s.chunk = chunk;
  s.per_period = (s.period + chunk - 1) / chunk;
  s.first_slot = (unsigned)(cursor / chunk);
  if (n == 0)
  {
    s.nchunks = 0;
  }
  else
  {
    const unsigned long long last = cursor + n - 1;
    const unsigned long long lp = last / s.period;
    const unsigned lr = (unsigned)((last % s.period) / chunk);
    s.nchunks = (unsigned)(lp * s.per_period + lr - s.first_slot + 1);
  }
  return s;
}

__host__ __device__ inline ChunkSpan chunk_span(const Schedule& s, unsigned j)
{
  const unsigned long long g = (unsigned long long)s.first_slot + j;
  const unsigned long long p = g / s.per_period;
  const unsigned r = (unsigned)(g - p * s.per_period);
  const unsigned long long base = p * s.period;
  unsigned long long us = base + (unsigned long long)r * s.chunk;
  unsigned long long ue = us + s.chunk;
  const unsigned long long pe = base + s.period;
  if (ue > pe) ue = pe;
  const unsigned long long call_end = s.cursor + s.n;
  if (us < s.cursor) us = s.cursor;
  if (ue > call_end) ue = call_end;
  ChunkSpan c;
  c.t0 = us - s.cursor;
  c.len = (unsigned)(ue - us);
  c.cursor0 = (unsigned)(us - base);
  c.first = (j == 0);
  c.wraps = (ue == pe);
  return c;
}

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
