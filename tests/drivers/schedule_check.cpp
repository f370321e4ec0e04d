// CPU check of the chunk schedule (sdft_b200/csrc/sdft_chunks.hpp): for many (cursor, n, m, L) the chunks must tile
// the call exactly and in order, never cross a multiple of L inside the period nor the period end, start on the
// L-grid (except the first chunk of a call), and flag the period's last step.  Exit code 0 = all invariants hold.
#include <cstdio>
#include <cstdlib>
#include "sdft_chunks.hpp"

using namespace sdftb200;

static unsigned long long rng_state = 0x9E3779B97F4A7C15ull;
static unsigned long long rnd()
{
  rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17;
  return rng_state;
}

int main()
{
  const unsigned chunks[] = { 32, 64, 96, 128, 256, 512, 1024 };
  long cases = 0;
  for (int it = 0; it < 200000; ++it)
  {
    const unsigned m = 1 + (unsigned)(rnd() % ((it % 3 == 0) ? 5000 : 70));
    const unsigned L = chunks[rnd() % 7];
    const unsigned period = 2 * m;
    const unsigned long long cursor = rnd() % period;
    const unsigned long long n = 1 + rnd() % ((it % 5 == 0) ? 5 * period + 7 : 3 * L + 5);
    const Schedule s = make_schedule(cursor, n, m, L);
    unsigned long long t = 0;
    for (unsigned j = 0; j < s.nchunks; ++j)
    {
      const ChunkSpan c = chunk_span(s, j);
      const unsigned long long abs0 = cursor + c.t0;            // position counted from the period start of the call
      const unsigned c0 = (unsigned)(abs0 % period);
      bool ok = (c.t0 == t) && (c.len >= 1) && (c.len <= L) && (c.cursor0 == c0) && (c.first == (j == 0));
      ok = ok && (c0 + c.len <= period);                         // never across the period end
      ok = ok && (c0 / L == (c0 + c.len - 1) / L);               // never across a multiple of L inside the period
      ok = ok && (j == 0 || c0 % L == 0);                        // on the L grid, except the call's first chunk
      ok = ok && (c.wraps == (c0 + c.len == period));            // last step of the period flagged
      if (!ok)
      {
        printf("schedule violation: m=%u L=%u cursor=%llu n=%llu chunk %u: t0=%llu len=%u cursor0=%u wraps=%d\n",
               m, L, cursor, n, j, c.t0, c.len, c.cursor0, (int)c.wraps);
        return 1;
      }
      t += c.len;
    }
    if (t != n || make_schedule(cursor, 0, m, L).nchunks != 0)
    {
      printf("schedule does not cover the call: m=%u L=%u cursor=%llu n=%llu covered=%llu\n", m, L, cursor, n, t);
      return 1;
    }
    ++cases;
  }
  // ticket interleaving of a mixed launch: a bijection onto (body tickets, tail tickets), each set in increasing order
  long mixes = 0;
  for (int it = 0; it < 3000; ++it)
  {
    const unsigned tail = 1 + (unsigned)(rnd() % 200), body = (unsigned)(rnd() % 4000) + (it % 7 == 0 ? 0 : tail);
    const unsigned every = mixed_every(body, tail);
    unsigned next_body = 0, next_tail = 0;
    for (unsigned t = 0; t < body + tail; ++t)
    {
      const MixedTicket mt = mixed_ticket(t, every, tail);
      unsigned& expect = mt.is_tail ? next_tail : next_body;
      if (every == 0 || mt.local != expect || (mt.is_tail ? mt.local >= tail : mt.local >= body))
      {
        printf("mixed ticket violation: body=%u tail=%u every=%u t=%u -> %s %u (expected %u)\n", body, tail, every, t,
               mt.is_tail ? "tail" : "body", mt.local, expect);
        return 1;
      }
      ++expect;
    }
    if (next_body != body || next_tail != tail)
    {
      printf("mixed tickets do not cover both sets: body=%u/%u tail=%u/%u every=%u\n", next_body, body, next_tail, tail, every);
      return 1;
    }
    ++mixes;
  }
  printf("%ld schedules ok, %ld ticket interleavings ok\n", cases, mixes);
  return 0;
}
