/*
 * sdft_weights.hpp -- host-side layout helpers that depend on the reference's mirror-cell rule (c/src/sdft/sdft.h:589-595):
 * which bins the four mirror cells copy, and the synthesis weights with the window folded in (the adjoint of
 * sdft_etc_convolve, c/src/sdft/sdft.h:350-402) that the fused analysis+synthesis kernel uses.  Plain C++; also
 * compiled on its own by the CPU tests (tests/test_weights.py).
 */
#pragma once

#include <cstddef>
#include <vector>

namespace sdftb200
{

/* mirror cells (sdft.h:589-595): source bin (or -1 = always zero) and whether the copy is conjugated;
 * only used on the host to lay out the extended twiddle table */
struct MirrorMap
{
  int cell[4];
  int src[4];
  int conj[4];
};

/* mirror cells resolved from the assignment order of sdft.h:589-595 */
inline MirrorMap make_mirrors(size_t m)
{
  MirrorMap mm;
  mm.cell[0] = 0; mm.cell[1] = 1; mm.cell[2] = (int)m + 2; mm.cell[3] = (int)m + 3;
  if (m >= 3)
  {
    mm.src[0] = 2; mm.conj[0] = 1;
    mm.src[1] = 1; mm.conj[1] = 1;
    mm.src[2] = (int)m - 2; mm.conj[2] = 1;
    mm.src[3] = (int)m - 3; mm.conj[3] = 1;
  }
  else if (m == 2)
  {
    /* aux[1]=conj(bin1); aux[4]=conj(bin0); aux[0]=conj(aux[4])=bin0; aux[5]=conj(aux[1])=bin1 */
    mm.src[0] = 0; mm.conj[0] = 0;
    mm.src[1] = 1; mm.conj[1] = 1;
    mm.src[2] = 0; mm.conj[2] = 1;
    mm.src[3] = 1; mm.conj[3] = 0;
  }
  else
  {
    /* m == 1: each mirror cell only ever copies itself through its partner and stays zero */
    for (int q = 0; q < 4; ++q) { mm.src[q] = -1; mm.conj[q] = 0; }
  }
  return mm;
}

/* For v[k] the per-bin factor of sdft_isdft (sdft.h:639-652) and T[j] the window taps,
 *     sum_k Re(v[k] * sum_j T[j] aux[k + j])  =  sum_b (A[b] Re(aux[b]) + B[b] Im(aux[b]))
 * with the mirror cells folded onto their source bins.  `ab` receives (A, B) per bin divided by `prescale` (the
 * factor the double fast mode folds into its spectra); returns true when every B is zero.  The tap constants are
 * rounded to F first, as the row path has them (sdft.h:366-393). */
template <typename F>
bool synth_weights(size_t dftsize, int window, const MirrorMap& mirrors, double prescale, const double* vr, const double* vi,
                   std::vector<F>& ab)
{
  const long m = (long)dftsize;
  const double w = (double)((F)(1) / (F)(dftsize * 2));
  double taps[5] = { 0, 0, w, 0, 0 };                       // T[-2 .. +2]
  if (window == 1) { taps[2] = 0.5 * w; taps[1] = taps[3] = -0.25 * w; }
  if (window == 2) { taps[2] = (double)(F)0.54 * w; taps[1] = taps[3] = -(double)(F)0.23 * w; }
  if (window == 3) { taps[2] = (double)(F)0.42 * w; taps[1] = taps[3] = -(double)(F)0.25 * w; taps[0] = taps[4] = (double)(F)0.04 * w; }
  std::vector<double> A(m, 0.0), B(m, 0.0);
  for (long k = 0; k < m; ++k)
  {
    for (int j = -2; j <= 2; ++j)
    {
      const double t = taps[j + 2];
      if (t == 0.0) continue;
      long cell = k + 2 + j, bin = cell - 2;
      bool conj = false;
      if (cell < 2 || cell >= m + 2)
      {
        int q = -1;
        for (int i = 0; i < 4; ++i)
          if (mirrors.cell[i] == (int)cell) q = i;
        if (q < 0 || mirrors.src[q] < 0) continue;         // a cell that is always zero
        bin = mirrors.src[q];
        conj = mirrors.conj[q] != 0;
      }
      /* Re(v a) = vr ar - vi ai ;  Re(v conj(a)) = vr ar + vi ai */
      A[bin] += t * vr[k];
      B[bin] += (conj ? +t : -t) * vi[k];
    }
  }
  const double unscale = 1.0 / prescale;
  ab.resize(2 * (size_t)m);
  bool unit = true;
  for (long b = 0; b < m; ++b)
  {
    ab[2 * b] = (F)(A[b] * unscale);
    ab[2 * b + 1] = (F)(B[b] * unscale);
    if (ab[2 * b + 1] != (F)0) unit = false;
  }
  return unit;
}

}  // namespace sdftb200
