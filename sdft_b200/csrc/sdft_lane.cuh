/*
 * sdft_lane.cuh -- per-lane engines of phase C: replay + window + row stores (EmitLane) and the fused synthesis (SynthLane) (c/src/sdft/sdft.h:570-597, :635-657).
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include "sdft_arith.cuh"

namespace sdftb200
{

/* ------------------------------------------------------------------------------------------------
 * Lane engine of the emit phase: replay a chunk from its carry, demodulate, apply the window across
 *     neighbouring cells and stream the (n, m) rows out.
 *
 *     One warp owns Geo<F, GEO>::WC consecutive cells (CPL per lane: 128 cells for double, 256 for float) of
 *     one chunk and is independent of every other warp: the outermost GROUP cells on either side are
 *     halo (recomputed by the neighbouring warp), so 124 (double) / 248 (float) bins per warp are
 *     stored; the boxcar window needs no halo.  Neighbour cells inside the warp come from registers
 *     or one shuffle.  Rows are written with consecutive lanes on consecutive bins, each lane storing
 *     32-byte groups (2 double or 4 float bins) with an evict-first policy when the row pitch allows
 *     it, else bin by bin.
 * ---------------------------------------------------------------------------------------------- */
/* cache policy of the row stores: rows are written once and not read again by this kernel */
#ifndef SDFT_B200_STORE_POLICY
#define SDFT_B200_STORE_POLICY ".L1::no_allocate.L2::evict_first"
#endif
/* one 32-byte group: 2 double bins or 4 float bins */
__device__ __forceinline__ void store_group(cx<double>* dst, const cx<double>* y)
{
  asm volatile("st.global" SDFT_B200_STORE_POLICY ".v4.f64 [%0], {%1, %2, %3, %4};"
               :: "l"(dst), "d"(y[0].r), "d"(y[0].i), "d"(y[1].r), "d"(y[1].i));
}
__device__ __forceinline__ void store_group(cx<float>* dst, const cx<float>* y)
{
  asm volatile("st.global" SDFT_B200_STORE_POLICY ".v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :: "l"(dst), "f"(y[0].r), "f"(y[0].i), "f"(y[1].r), "f"(y[1].i),
                  "f"(y[2].r), "f"(y[2].i), "f"(y[3].r), "f"(y[3].i));
}
__device__ __forceinline__ void store_one(cx<double>* dst, cx<double> a)
{
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};" :: "l"(dst), "d"(a.r), "d"(a.i));
}
__device__ __forceinline__ void store_one(cx<float>* dst, cx<float> a)
{
  asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};" :: "l"(dst), "f"(a.r), "f"(a.i));
}

template <typename F>
__device__ __forceinline__ cx<F> shfl_up1(cx<F> v)
{
  cx<F> o;
  o.r = __shfl_up_sync(0xffffffffu, v.r, 1);
  o.i = __shfl_up_sync(0xffffffffu, v.i, 1);
  return o;
}
template <typename F>
__device__ __forceinline__ cx<F> shfl_down1(cx<F> v)
{
  cx<F> o;
  o.r = __shfl_down_sync(0xffffffffu, v.r, 1);
  o.i = __shfl_down_sync(0xffffffffu, v.i, 1);
  return o;
}

/* accumulate / demodulate / window stages of the modulated replay with the reference's own roundings */
template <typename F, bool FUSED> struct StageOps
{
  static __device__ __forceinline__ cx<F> mac(cx<F> acc, cx<F> p, F d) { return Arith<F>::mac(acc, p, d); }
  static __device__ __forceinline__ cx<F> demod(cx<F> a, cx<F> p) { return Arith<F>::demod(a, p); }
  template <int WINDOW>
  static __device__ __forceinline__ cx<F> window(cx<F> l2, cx<F> l1, cx<F> c, cx<F> r1, cx<F> r2, const WindowConst<F>& k)
  {
    return Arith<F>::template window<WINDOW>(l2, l1, c, r1, r2, k);
  }
};
template <typename F, int WINDOW, int GEO> struct EmitGeo
{
  enum
  {
    CPL = Geo<F, GEO>::CPL,
    GROUP = Geo<F, GEO>::GROUP,
    NGROUP = CPL / GROUP,
    WC = Geo<F, GEO>::WC,
    HALO = (WINDOW == 0) ? 0 : (int)GROUP,
    SPAN = WC - 2 * HALO        // bins stored per warp
  };
};

template <typename F, int WINDOW, bool VEC, int GEO>
struct EmitLane
{
  typedef EmitGeo<F, WINDOW, GEO> G;
  cx<F> acc[G::CPL];
  cx<F> ph[G::CPL];
  cx<F> tw[G::CPL];
  cx<F>* dst;            // address of this lane's cell 0 in the current row (may be out of range)
  bool ok[G::CPL];

  /* geometry of lane `lane` of warp-group `group`: first cell index (signed: the float halo reaches
   * below cell 0) and which of its cells are stored */
  __device__ __forceinline__ int setup(unsigned group, unsigned lane, unsigned m, unsigned roi_first, unsigned roi_end,
                                       unsigned bin_base, unsigned bin_end)
  {
    /* this launch covers bins [bin_base, bin_end) -- all of them, unless the call is split into a wide body and
     * a narrow tail (sdft_launch.hpp) */
    const int e0 = (int)(bin_base + group * G::SPAN) + 2 - G::HALO + (int)(lane * G::CPL);
#pragma unroll
    for (int b = 0; b < G::CPL; ++b)
    {
      const int slot = (int)lane * G::CPL + b;
      const int e = e0 + b;
      ok[b] = (slot >= G::HALO) && (slot < G::WC - G::HALO) && (e >= 2) && (e < (int)m + 2) && (e < (int)bin_end + 2) &&
              (e >= (int)roi_first + 2) && (e < (int)roi_end + 2);      // bins outside the region of interest are not stored
    }
    return e0;
  }

  /* one time step; RESTART = the period's last step, after which the phase restarts (sdft.h:566-576) */
  __device__ __forceinline__ void store_rows(const cx<F>* y, size_t row_stride)
  {
    if (VEC)
    {
#pragma unroll
      for (int g = 0; g < G::NGROUP; ++g)
        if (ok[g * G::GROUP]) store_group(dst + g * G::GROUP, y + g * G::GROUP);
    }
    else
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
        if (ok[b]) store_one(dst + b, y[b]);
    }
    dst += row_stride;
  }

  template <bool RESTART, bool FUSED>
  __device__ __forceinline__ void step(F d, const cx<F>* restart, const WindowConst<F>& win, size_t row_stride)
  {
    cx<F> y[G::CPL];
    compute<RESTART, FUSED>(d, restart, win, y);
    store_rows(y, row_stride);
  }

  __device__ __forceinline__ void fast_step(F d, const WindowConst<F>& win, size_t row_stride)
  {
    cx<F> y[G::CPL];
    fast_compute(d, win, y);
    store_rows(y, row_stride);
  }

  /* one time step of the modulated replay WITHOUT the window: the demodulated spectrum of this lane's
   * cells (sdft.h:583-585) into x[] */
  template <bool RESTART, bool FUSED>
  __device__ __forceinline__ void advance(F d, const cx<F>* restart, cx<F>* x)
  {
    typedef Arith<F> A;
    typedef StageOps<F, FUSED> S;
#pragma unroll
    for (int b = 0; b < G::CPL; ++b)
    {
      acc[b] = S::mac(acc[b], ph[b], d);
      ph[b] = RESTART ? restart[b] : A::rotate(ph[b], tw[b]);
      x[b] = S::demod(acc[b], ph[b]);
    }
  }

  /* one time step of the modulated replay: windowed spectrum of this lane's cells into y[] */
  template <bool RESTART, bool FUSED>
  __device__ __forceinline__ void compute(F d, const cx<F>* restart, const WindowConst<F>& win, cx<F>* y)
  {
    typedef StageOps<F, FUSED> S;
    cx<F> x[G::CPL];
    advance<RESTART, FUSED>(d, restart, x);
    if (WINDOW == 0)
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) y[b] = S::template window<0>(x[b], x[b], x[b], x[b], x[b], win);
    }
    else
    {
      /* neighbours: [l2 l1 | x0 .. x(CPL-1) | r1 r2] */
      const cx<F> l1 = shfl_up1(x[G::CPL - 1]);
      const cx<F> r1 = shfl_down1(x[0]);
      cx<F> l2 = l1, r2 = r1;   // only read by the 5-tap window
      if (WINDOW == 3)
      {
        l2 = shfl_up1(x[G::CPL - 2]);
        r2 = shfl_down1(x[1]);
      }
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        const cx<F> m2 = (b >= 2) ? x[b >= 2 ? b - 2 : 0] : ((b == 1) ? l1 : l2);
        const cx<F> m1 = (b >= 1) ? x[b >= 1 ? b - 1 : 0] : l1;
        const cx<F> p1 = (b + 1 < G::CPL) ? x[b + 1 < G::CPL ? b + 1 : 0] : r1;
        const cx<F> p2 = (b + 2 < G::CPL) ? x[b + 2 < G::CPL ? b + 2 : 0] : ((b + 1 < G::CPL) ? r1 : r2);
        y[b] = S::template window<WINDOW>(m2, m1, x[b], p1, p2, win);
      }
    }
  }

  /* fast mode (double): tw[] holds conj(tw), ph[] is unused, and acc[] holds z_t = aux_{t-1} + delta_t,
   * the demodulated spectrum BEFORE its rotation: aux_t = z_t conj(tw), so
   *     z_{t+1} = z_t conj(tw) + delta_{t+1}            one Horner step, 4 FP64 instructions,
   *     aux_t   = z_{t+1} - delta_{t+1}.
   * The window is linear and delta is the same real number in every cell (mirror cells included), so
   *     window(aux_t) = window(z_{t+1}) - delta_{t+1} * (sum of the taps):
   * nothing to subtract for hann and blackman (their taps sum to zero), one real subtraction per bin
   * for boxcar and hamming.  The caller passes d_next = delta_{t+1}, 0 after the chunk's last sample
   * (then z_{t+1} IS aux_t), and seeds z_0 = anchor + delta_0. */
  __device__ __forceinline__ void fast_compute(F d_next, const WindowConst<F>& win, cx<F>* y)
  {
    typedef Arith<F> A;
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) acc[b] = A::horner(acc[b], tw[b], d_next);
    fast_output(acc, d_next, win, y);
  }

  /* software-pipelined form: acc[] already holds z_{t+1}; the recurrence for step t+1 (z_{t+2}, needs
   * d_after = delta_{t+2}) is issued FIRST so that its FP64 latency overlaps the shuffles, taps and
   * stores of step t.  Matters when few warps share an SM (short calls): the in-order issue would
   * otherwise expose every latency of a step before the next one starts. */
  __device__ __forceinline__ void fast_compute_ahead(F d_next, F d_after, const WindowConst<F>& win, cx<F>* y)
  {
    typedef Arith<F> A;
    cx<F> nxt[G::CPL];
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) nxt[b] = A::horner(acc[b], tw[b], d_after);
    fast_output(acc, d_next, win, y);
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) acc[b] = nxt[b];
  }

  /* window(z) - d_next * (sum of taps), see fast_compute */
  __device__ __forceinline__ void fast_output(const cx<F>* acc, F d_next, const WindowConst<F>& win, cx<F>* y)
  {
    typedef Arith<F> A;
    if (WINDOW == 0)
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        y[b].r = __dadd_rn(acc[b].r, -d_next);
        y[b].i = acc[b].i;
      }
    }
    else
    {
      const cx<F> l1 = shfl_up1(acc[G::CPL - 1]);
      const cx<F> r1 = shfl_down1(acc[0]);
      cx<F> l2 = l1, r2 = r1;
      if (WINDOW == 3)
      {
        l2 = shfl_up1(acc[G::CPL - 2]);
        r2 = shfl_down1(acc[1]);
      }
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        const cx<F> m2 = (b >= 2) ? acc[b >= 2 ? b - 2 : 0] : ((b == 1) ? l1 : l2);
        const cx<F> m1 = (b >= 1) ? acc[b >= 1 ? b - 1 : 0] : l1;
        const cx<F> p1 = (b + 1 < G::CPL) ? acc[b + 1 < G::CPL ? b + 1 : 0] : r1;
        const cx<F> p2 = (b + 2 < G::CPL) ? acc[b + 2 < G::CPL ? b + 2 : 0] : ((b + 1 < G::CPL) ? r1 : r2);
        y[b] = A::template fast_window<WINDOW>(m2, m1, acc[b], p1, p2, win);
      }
      if (WINDOW == 2)
      {
        const F corr = __dmul_rn(d_next, win.ksum);
#pragma unroll
        for (int b = 0; b < G::CPL; ++b) y[b].r = __dadd_rn(y[b].r, -corr);
      }
    }
  }
};

/* ------------------------------------------------------------------------------------------------
 * Fused synthesis (EMIT_SYNTH): instead of storing the rows, every lane weighs its bins and the warp
 * reduces over the bins.  The window never runs on the device here: sdft_isdft of a windowed row is
 *     y = 2 sum_k Re(v[k] * sum_j T[j] aux[k + j])          v = (-1)^k or tws[k]   (sdft.h:639-652),
 * linear in aux, so the taps T are moved onto the weights on the host (the adjoint of sdft_etc_convolve,
 * mirror cells folded onto their source bins, make_synth_weights):
 *     y = 2 sum_b (A[b] Re(aux[b]) + B[b] Im(aux[b])).
 * No halo, no shuffles across bins, 1-2 FMAs per bin instead of the 4-10 of the taps.  The warp reduces
 * eight time steps at once with a transposing butterfly (9 shuffles per 8 steps instead of 5 per step):
 * after three exchange rounds every lane holds ONE step's sum over eight lanes, two plain butterfly
 * rounds finish it.  The warp's partial sums go to part[group][t]; a second tiny kernel adds the groups in
 * order and scales by 2 (sdft.h:654-656).  Fixed order: deterministic.
 * ---------------------------------------------------------------------------------------------- */
template <typename F, int CPL, bool UNIT>
struct SynthLane
{
  F wa[CPL], wb[CPL];   // weights of Re / Im of this lane's bins (wb all zero when UNIT); 0 beyond bin m-1
  F wsum;               // sum of wa[]: what a constant added to every cell's real part contributes
  F p[8];

  __device__ __forceinline__ void setup(const F* __restrict__ ab, int e0, const bool* ok)
  {
    wsum = (F)0;
#pragma unroll
    for (int b = 0; b < CPL; ++b)
    {
      const int k = e0 + b - 2;
      wa[b] = ok[b] ? ab[2 * k] : (F)0;
      wb[b] = (ok[b] && !UNIT) ? ab[2 * k + 1] : (F)0;
      wsum += wa[b];
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = (F)0;
  }

  __device__ __forceinline__ F weigh(const cx<F>* x) const
  {
    F s = (F)0;
#pragma unroll
    for (int b = 0; b < CPL; ++b)
    {
      s = fma(x[b].r, wa[b], s);
      if (!UNIT) s = fma(x[b].i, wb[b], s);
    }
    return s;
  }

  /* sums p[0..7] over the warp; lane (4 q) returns the total of step q's slot, see step_of() */
  __device__ __forceinline__ F reduce8(unsigned lane)
  {
    F a[4], b2[2], c;
    {
      const bool hi = (lane & 16) != 0;
#pragma unroll
      for (int i = 0; i < 4; ++i)
      {
        const F keep = hi ? p[4 + i] : p[i];
        const F give = hi ? p[i] : p[4 + i];
        a[i] = keep + __shfl_xor_sync(0xffffffffu, give, 16);
      }
    }
    {
      const bool hi = (lane & 8) != 0;
#pragma unroll
      for (int i = 0; i < 2; ++i)
      {
        const F keep = hi ? a[2 + i] : a[i];
        const F give = hi ? a[i] : a[2 + i];
        b2[i] = keep + __shfl_xor_sync(0xffffffffu, give, 8);
      }
    }
    {
      const bool hi = (lane & 4) != 0;
      const F keep = hi ? b2[1] : b2[0];
      const F give = hi ? b2[0] : b2[1];
      c = keep + __shfl_xor_sync(0xffffffffu, give, 4);
    }
    c += __shfl_xor_sync(0xffffffffu, c, 2);
    c += __shfl_xor_sync(0xffffffffu, c, 1);
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = (F)0;
    return c;
  }
  static __device__ __forceinline__ F warp_sum(F v)
  {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
  }
  /* which of the eight steps lane `lane` holds after reduce8 */
  static __device__ __forceinline__ unsigned step_of(unsigned lane)
  {
    return ((lane >> 4) & 1u) * 4u + ((lane >> 3) & 1u) * 2u + ((lane >> 2) & 1u);
  }
};

/* part: (channels, groups, n) partial sums -> samples (channels, sample_stride), sdft.h:654-656 */
template <typename T, typename F>
__global__ void synth_finish_kernel(const F* __restrict__ part, unsigned groups, unsigned long long n,
                                    T* __restrict__ samples, size_t sample_stride)
{
  const unsigned ch = blockIdx.y;
  const F* base = part + (size_t)ch * groups * n;
  T* y = samples + (size_t)ch * sample_stride;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
  {
    F s = (F)0;
    for (unsigned g = 0; g < groups; ++g) s += base[(size_t)g * n + t];
    y[t] = (T)(s * (F)2);
  }
}



}  // namespace sdftb200
