/*
 * sdft_tables.hpp -- analysis and synthesis twiddle tables in the reference's exact expression order and types
 * (c/src/sdft/sdft.h:439-446), computed with the HOST libm so that float tables are bit-identical to the
 * reference's; the device never evaluates sin/cos.  Plain C++; also compiled on its own by the
 * CPU tests (tests/test_tables.py), which compare the tables bit for bit with the reference's.
 */
#pragma once

#include <cmath>
#include <cstddef>
#include <vector>

namespace sdftb200
{

template <typename F> struct table_entry { F r, i; };   // layout-identical to cx<F>

/* ------------------------------------------------------------------------------------------------
 * host-side trigonometry with the reference's expression order and types (sdft.h:439-446).
 * Computed with the host libm so that float tables are bit-identical to the reference's
 * (SURVEY.md fact 5); the device never evaluates sin/cos.
 * ---------------------------------------------------------------------------------------------- */
inline float t_cos(float x) { return ::cosf(x); }
inline float t_sin(float x) { return ::sinf(x); }
inline float t_acos(float x) { return ::acosf(x); }
inline double t_cos(double x) { return ::cos(x); }
inline double t_sin(double x) { return ::sin(x); }
inline double t_acos(double x) { return ::acos(x); }

template <typename F, typename E = table_entry<F>>
void make_tables(size_t m, double latency, std::vector<E>& tw, std::vector<E>& tws)
{
  tw.resize(m);
  tws.resize(m);
  const F omega = (F)(-2) * t_acos((F)(-1)) / (F)(m * 2);
  const F wsyn = (F)(+2) / ((F)(1) - t_cos((F)((omega * (F)m) * latency)));
  for (size_t k = 0; k < m; ++k)
  {
    const F a = omega * (F)k;
    tw[k].r = (F)(1) * t_cos(a);
    tw[k].i = (F)(1) * t_sin(a);
    const F s = (F)(((omega * (F)k) * (F)m) * latency);   // trailing product in double, then narrowed
    tws[k].r = wsyn * t_cos(s);
    tws[k].i = wsyn * t_sin(s);
  }
}

}  // namespace sdftb200
