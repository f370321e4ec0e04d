/*
 * sdft_chunks.hpp -- the chunk schedule of one analysis call: how [cursor, cursor + n) is cut into chunks that
 * never cross a multiple of the chunk length inside the 2m period nor the period end (where the reference
 * restarts its modulation phase, c/src/sdft/sdft.h:566-576).  Plain C++, identical code on host and device;
 * also compiled on its own by the CPU tests (tests/test_schedule.py).
 */
#pragma once

#if defined(__CUDACC__)
#define SDFT_B200_HD __host__ __device__
#else
#define SDFT_B200_HD
#endif

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

SDFT_B200_HD inline Schedule make_schedule(unsigned long long cursor, unsigned long long n,
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

SDFT_B200_HD inline ChunkSpan chunk_span(const Schedule& s, unsigned j)
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
 * ticket interleaving of a mixed launch (scan_emit_mixed_kernel): `body` + `tail` work tickets are handed out
 * by ONE counter; every `every`-th ticket (indices every-1, 2*every-1, ...) belongs to the tail set until the tail
 * has `tail` of them, all others to the body set.  Both sets see their own tickets in increasing order, so a CTA
 * only ever waits for CTAs with smaller global tickets.  every = (body + tail) / tail keeps the last tail
 * ticket inside the launch.
 * ---------------------------------------------------------------------------------------------- */
struct MixedTicket
{
  bool is_tail;
  unsigned local;     // ticket inside its own set
};

SDFT_B200_HD inline unsigned mixed_every(unsigned body, unsigned tail) { return (body + tail) / tail; }

SDFT_B200_HD inline MixedTicket mixed_ticket(unsigned t, unsigned every, unsigned tail)
{
  const unsigned slot = t / every;               // tail tickets handed out before t (if the tail still had some)
  MixedTicket r;
  r.is_tail = ((t + 1u) % every == 0u) && (slot < tail);
  r.local = r.is_tail ? slot : t - (slot < tail ? slot : tail);
  return r;
}

}  // namespace sdftb200
