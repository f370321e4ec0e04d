/*
 * sdft_kernels.cuh -- sm_100a kernels of the sliding-DFT hot path.
 *
 * Reference being replaced: the per-sample loops of c/src/sdft/sdft.h:562-598 (analysis),
 * :350-402 (window convolution) and :635-672 (synthesis).  See DESIGN.md for the scan formulation.
 *
 * Vocabulary
 *   bin k            0..m-1, the reference's dft index
 *   cell e = k + 2   "extended" index 0..m+3; cells 0,1,m+2,m+3 are the mirror cells of
 *                    sdft.h:589-595.  They are carried as ordinary bins with conjugated twiddles:
 *                    (a+bi)(c+di) and its conjugate round identically, so a mirror cell evolves as the
 *                    exact conjugate of its source bin and no mirroring step exists on the device.
 *   phase P[c][e]    the reference's "fiddle": tw^c by sequential multiplication, restarted every
 *                    2m samples (sdft.h:566-576)
 *   chunk            a run of <= L consecutive samples that never crosses a multiple of L inside the
 *                    2m period nor the period end; every chunk but the first of a call therefore
 *                    starts at a cursor that is a multiple of L and reads its phase from the F0 table
 *
 * Arithmetic policy
 *   float   reproduces every rounding point of the reference (un-fused, packed FMUL2/FADD2): phases
 *           are bit-exact and rows are bit-exact within a chunk; only the summation order across
 *           chunks differs.
 *   double  MODE_MODULATED: the reference's scheme with explicit FMAs in a fixed pattern.
 *           MODE_FAST (default): the chunk total is a Horner sum of tw^i * delta_i scaled by the phase
 *           at the chunk start, and the replay runs the demodulated recurrence
 *           aux <- (aux + delta) * conj(tw) anchored at carry * conj(P_start) at every chunk start,
 *           with the window weight folded into the deltas.  13 instead of 22 FP64 instructions per
 *           bin-update (Hann); the B200 is power-capped on this path, so fewer FP64 operations is
 *           more bandwidth.  Mathematically identical, differs by rounding (~1e-12 of full scale,
 *           gate is 1e-9); the phase still restarts exactly every 2m samples.
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sdftb200
{

constexpr int kF0Stride = 32;     // phase table holds P at every 32nd cursor
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

/* ------------------------------------------------------------------------------------------------
 * arithmetic policies (complex level)
 * ---------------------------------------------------------------------------------------------- */
template <typename F> struct WindowConst
{
  F w;       // analysis weight 1/(2m)                        sdft.h:422
  F wq;      // w * 0.25, the Hann factor                     sdft.h:371
  F c0, c1, c2;   // double, modulated mode: weight folded into centre / first / second neighbour taps
  F pre;          // double, fast mode: factor folded into the deltas (whole weight times one tap)
  F k0, k1;       // double, fast mode: remaining tap ratios
  F ksum;         // double, fast mode: sum of the taps on pre-scaled data (0 for hann and blackman)
};

template <typename F>
inline WindowConst<F> make_window_const(size_t m, int window)
{
  WindowConst<F> k;
  k.w = (F)(1) / (F)(m * 2);
  k.wq = k.w * (F)(0.25);
  switch (window)
  {
    case 1: k.c0 = (F)2 * k.wq; k.c1 = k.wq; k.c2 = (F)0; break;
    case 2: k.c0 = (F)(0.54) * k.w; k.c1 = (F)(0.23) * k.w; k.c2 = (F)0; break;
    case 3: k.c0 = (F)(0.42) * k.w; k.c1 = (F)(0.25) * k.w; k.c2 = (F)(0.04) * k.w; break;
    default: k.c0 = k.w; k.c1 = (F)0; k.c2 = (F)0; break;
  }
  /* fast mode, taps on pre-scaled data: hann 2c-(l+r); hamming k0*c-(l+r);
   * blackman (l2+r2) + k0*c - k1*(l1+r1); boxcar c */
  switch (window)
  {
    case 1: k.pre = k.wq; k.k0 = (F)2; k.k1 = (F)0; break;
    case 2: k.pre = (F)(0.23) * k.w; k.k0 = (F)(0.54) / (F)(0.23); k.k1 = (F)0; break;
    case 3: k.pre = (F)(0.04) * k.w; k.k0 = (F)(0.42) / (F)(0.04); k.k1 = (F)(0.25) / (F)(0.04); break;
    default: k.pre = k.w; k.k0 = (F)1; k.k1 = (F)0; break;
  }
  k.ksum = (window == 0) ? (F)1 : ((window == 2) ? k.k0 - (F)2 : (F)0);
  return k;
}

template <typename F> struct Arith;

/* float: every operation of the reference is kept as its own rounding step (bit-exact phases and,
 * within a chunk, bit-exact rows).  Products use the packed FMUL2/FADD2 forms of sm_100a on the
 * (re, im) register pair to halve the issue slots.  ptxas contracts a packed mul.rn.f32x2 feeding a
 * packed add/sub.rn.f32x2 into FFMA2 (observed with CUDA 12.9, even with --fmad=false), so every
 * addition that consumes a product is a SCALAR add.rn/sub.rn, which ptxas never fuses. */
__device__ __forceinline__ cx<float> pk_mul(cx<float> a, cx<float> b)        // (a.r*b.r, a.i*b.i)
{
  cx<float> o;
  asm("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %5}; mul.rn.f32x2 z, x, y; mov.b64 {%0, %1}, z;}"
      : "=f"(o.r), "=f"(o.i) : "f"(a.r), "f"(a.i), "f"(b.r), "f"(b.i));
  return o;
}
__device__ __forceinline__ cx<float> pk_mul_cross(cx<float> a, cx<float> b)  // (a.r*b.i, a.i*b.r)
{
  cx<float> o;
  asm("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%5, %4}; mul.rn.f32x2 z, x, y; mov.b64 {%0, %1}, z;}"
      : "=f"(o.r), "=f"(o.i) : "f"(a.r), "f"(a.i), "f"(b.r), "f"(b.i));
  return o;
}
__device__ __forceinline__ cx<float> pk_scale(cx<float> a, float k)          // (a.r*k, a.i*k)
{
  cx<float> o;
  asm("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mul.rn.f32x2 z, x, y; mov.b64 {%0, %1}, z;}"
      : "=f"(o.r), "=f"(o.i) : "f"(a.r), "f"(a.i), "f"(k));
  return o;
}
__device__ __forceinline__ cx<float> pk_add(cx<float> a, cx<float> b)
{
  cx<float> o;
  asm("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %5}; add.rn.f32x2 z, x, y; mov.b64 {%0, %1}, z;}"
      : "=f"(o.r), "=f"(o.i) : "f"(a.r), "f"(a.i), "f"(b.r), "f"(b.i));
  return o;
}
__device__ __forceinline__ cx<float> pk_sub(cx<float> a, cx<float> b)
{
  cx<float> o;
  asm("{.reg .b64 x, y, z; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %5}; sub.rn.f32x2 z, x, y; mov.b64 {%0, %1}, z;}"
      : "=f"(o.r), "=f"(o.i) : "f"(a.r), "f"(a.i), "f"(b.r), "f"(b.i));
  return o;
}

__device__ __forceinline__ cx<float> pk_fma(cx<float> a, cx<float> b, cx<float> c)   // (a.r*b.r+c.r, a.i*b.i+c.i)
{
  cx<float> o;
  asm("{.reg .b64 x, y, z, w; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %5}; mov.b64 z, {%6, %7}; fma.rn.f32x2 w, x, y, z; mov.b64 {%0, %1}, w;}"
      : "=f"(o.r), "=f"(o.i) : "f"(a.r), "f"(a.i), "f"(b.r), "f"(b.i), "f"(c.r), "f"(c.i));
  return o;
}
__device__ __forceinline__ cx<float> pk_fma_s(cx<float> a, float k, cx<float> c)           // (a.r*k+c.r, a.i*k+c.i)
{
  cx<float> o;
  asm("{.reg .b64 x, y, z, w; mov.b64 x, {%2, %3}; mov.b64 y, {%4, %4}; mov.b64 z, {%5, %6}; fma.rn.f32x2 w, x, y, z; mov.b64 {%0, %1}, w;}"
      : "=f"(o.r), "=f"(o.i) : "f"(a.r), "f"(a.i), "f"(k), "f"(c.r), "f"(c.i));
  return o;
}

template <> struct Arith<float>
{
  typedef float F;
  /* ---- fused variants (MODE_FAST): only used where the result is NOT fed back into the modulation
   * phase.  The phase recurrence `rotate` stays un-fused in every mode: its rounding compounds over up
   * to 2m-1 steps and must be the reference's bit for bit; a fused accumulate or a fused
   * output stage moves a value by <= 1 ulp, the same order as the chunked summation order does. ---- */
  static __device__ __forceinline__ cx<F> mac_fused(cx<F> acc, cx<F> p, F d) { return pk_fma_s(p, d, acc); }
  static __device__ __forceinline__ cx<F> demod_fused(cx<F> a, cx<F> p)
  {
    /* (ar*pr + ai*pi, ai*pr - ar*pi) */
    cx<F> sw, np;
    sw.r = a.i; sw.i = a.r;
    np.r = p.i; np.i = -p.i;
    const cx<F> u = pk_mul(sw, np);          // (ai*pi, -ar*pi)
    cx<F> pr;
    pr.r = p.r; pr.i = p.r;
    return pk_fma(a, pr, u);
  }
  template <int WINDOW>
  static __device__ __forceinline__ cx<F> window_fused(cx<F> l2, cx<F> l1, cx<F> c, cx<F> r1, cx<F> r2, const WindowConst<F>& k)
  {
    if (WINDOW == 0) return pk_scale(c, k.c0);
    const cx<F> s1 = pk_scale(pk_add(l1, r1), -k.c1);
    cx<F> y = pk_fma_s(c, k.c0, s1);
    if (WINDOW == 3) y = pk_fma_s(pk_add(l2, r2), k.c2, y);
    return y;
  }
  static __device__ __forceinline__ cx<F> cadd(cx<F> a, cx<F> b)
  {
    cx<F> o;
    o.r = __fadd_rn(a.r, b.r);
    o.i = __fadd_rn(a.i, b.i);
    return o;
  }
  /* P * tw, sdft.h:584 via :298-300 -- bit-exact with the reference */
  static __device__ __forceinline__ cx<F> rotate(cx<F> p, cx<F> w)
  {
    const cx<F> t1 = pk_mul(p, w);         // (pr*wr, pi*wi)
    const cx<F> t2 = pk_mul_cross(p, w);   // (pr*wi, pi*wr)
    cx<F> o;
    o.r = __fsub_rn(t1.r, t1.i);
    o.i = __fadd_rn(t2.r, t2.i);
    return o;
  }
  /* acc + P * delta, sdft.h:583 */
  static __device__ __forceinline__ cx<F> mac(cx<F> acc, cx<F> p, F d)
  {
    const cx<F> t = pk_scale(p, d);
    cx<F> o;
    o.r = __fadd_rn(acc.r, t.r);
    o.i = __fadd_rn(acc.i, t.i);
    return o;
  }
  /* acc * conj(P), sdft.h:585: (ar*pr - ai*(-pi), ar*(-pi) + ai*pr) */
  static __device__ __forceinline__ cx<F> demod(cx<F> a, cx<F> p)
  {
    const cx<F> t1 = pk_mul(a, p);         // (ar*pr, ai*pi)
    const cx<F> t2 = pk_mul_cross(a, p);   // (ar*pi, ai*pr)
    cx<F> o;
    o.r = __fadd_rn(t1.r, t1.i);
    o.i = __fsub_rn(t2.i, t2.r);
    return o;
  }
  /* window taps in the reference's operation order, sdft.h:350-402 */
  template <int WINDOW>
  static __device__ __forceinline__ cx<F> window(cx<F> l2, cx<F> l1, cx<F> c, cx<F> r1, cx<F> r2, const WindowConst<F>& k)
  {
    if (WINDOW == 1)
    {
      return pk_scale(pk_sub(pk_add(c, c), pk_add(l1, r1)), k.wq);
    }
    else if (WINDOW == 2)
    {
      const cx<F> a = pk_scale(c, (F)(0.54));
      const cx<F> b = pk_scale(pk_add(l1, r1), (F)(0.23));
      cx<F> d;
      d.r = __fsub_rn(a.r, b.r);
      d.i = __fsub_rn(a.i, b.i);
      return pk_scale(d, k.w);
    }
    else if (WINDOW == 3)
    {
      const cx<F> a = pk_scale(c, (F)(0.42));
      const cx<F> b = pk_scale(pk_add(l1, r1), (F)(0.25));
      const cx<F> e = pk_scale(pk_add(l2, r2), (F)(0.04));
      cx<F> d;
      d.r = __fadd_rn(__fsub_rn(a.r, b.r), e.r);
      d.i = __fadd_rn(__fsub_rn(a.i, b.i), e.i);
      return pk_scale(d, k.w);
    }
    else
    {
      return pk_scale(c, k.w);
    }
  }
};

/* double: explicit FMAs in a FIXED pattern (every kernel generates bit-identical phases); the window
 * weight is folded into the tap coefficients (3 / 3 / 5 FP64 instructions per component instead of
 * 4 / 5 / 8).  Differs from the reference's operation order by rounding only (~1e-16). */
template <> struct Arith<double>
{
  typedef double F;
  static __device__ __forceinline__ cx<F> cadd(cx<F> a, cx<F> b)
  {
    cx<F> o;
    o.r = __dadd_rn(a.r, b.r);
    o.i = __dadd_rn(a.i, b.i);
    return o;
  }
  static __device__ __forceinline__ cx<F> rotate(cx<F> p, cx<F> w)
  {
    cx<F> o;
    o.r = __fma_rn(p.r, w.r, -__dmul_rn(p.i, w.i));
    o.i = __fma_rn(p.r, w.i, __dmul_rn(p.i, w.r));
    return o;
  }
  static __device__ __forceinline__ cx<F> mac(cx<F> acc, cx<F> p, F d)
  {
    cx<F> o;
    o.r = __fma_rn(p.r, d, acc.r);
    o.i = __fma_rn(p.i, d, acc.i);
    return o;
  }
  static __device__ __forceinline__ cx<F> demod(cx<F> a, cx<F> p)
  {
    cx<F> o;
    o.r = __fma_rn(a.r, p.r, __dmul_rn(a.i, p.i));
    o.i = __fma_rn(a.i, p.r, -__dmul_rn(a.r, p.i));
    return o;
  }
  template <int WINDOW>
  static __device__ __forceinline__ F tap(F l2, F l1, F c, F r1, F r2, const WindowConst<F>& k)
  {
    if (WINDOW == 0)
    {
      return __dmul_rn(c, k.c0);
    }
    else if (WINDOW == 3)
    {
      const F t = __fma_rn(c, k.c0, -__dmul_rn(__dadd_rn(l1, r1), k.c1));
      return __fma_rn(__dadd_rn(l2, r2), k.c2, t);
    }
    else
    {
      return __fma_rn(c, k.c0, -__dmul_rn(__dadd_rn(l1, r1), k.c1));
    }
  }
  template <int WINDOW>
  static __device__ __forceinline__ cx<F> window(cx<F> l2, cx<F> l1, cx<F> c, cx<F> r1, cx<F> r2, const WindowConst<F>& k)
  {
    cx<F> o;
    o.r = tap<WINDOW>(l2.r, l1.r, c.r, r1.r, r2.r, k);
    o.i = tap<WINDOW>(l2.i, l1.i, c.i, r1.i, r2.i, k);
    return o;
  }

  /* ---- fast mode ---- */
  /* Horner step of the chunk total: h <- h * tw + delta */
  static __device__ __forceinline__ cx<F> horner(cx<F> h, cx<F> w, F d)
  {
    cx<F> o;
    o.r = __fma_rn(h.r, w.r, __fma_rn(-h.i, w.i, d));
    o.i = __fma_rn(h.r, w.i, __dmul_rn(h.i, w.r));
    return o;
  }
  /* four Horner steps at once: h <- h * tw^4 + (d0 + d1 tw + d2 tw^2 + d3 tw^3); 10 instead of 16
   * FP64 instructions */
  static __device__ __forceinline__ cx<F> horner4(cx<F> h, cx<F> w1, cx<F> w2, cx<F> w3, cx<F> w4, F d0, F d1, F d2, F d3)
  {
    const F ir = __fma_rn(d3, w3.r, __fma_rn(d2, w2.r, __fma_rn(d1, w1.r, d0)));
    const F ii = __fma_rn(d3, w3.i, __fma_rn(d2, w2.i, __dmul_rn(d1, w1.i)));
    cx<F> o;
    o.r = __fma_rn(h.r, w4.r, __fma_rn(-h.i, w4.i, ir));
    o.i = __fma_rn(h.r, w4.i, __fma_rn(h.i, w4.r, ii));
    return o;
  }
  static __device__ __forceinline__ cx<F> cmul(cx<F> a, cx<F> b)
  {
    cx<F> o;
    o.r = __fma_rn(a.r, b.r, -__dmul_rn(a.i, b.i));
    o.i = __fma_rn(a.r, b.i, __dmul_rn(a.i, b.r));
    return o;
  }
  /* demodulated recurrence: aux <- (aux + delta) * cw with cw = conj(tw) */
  static __device__ __forceinline__ cx<F> slide(cx<F> x, cx<F> cw, F d)
  {
    const F t = __dadd_rn(x.r, d);
    cx<F> o;
    o.r = __fma_rn(t, cw.r, -__dmul_rn(x.i, cw.i));
    o.i = __fma_rn(t, cw.i, __dmul_rn(x.i, cw.r));
    return o;
  }
  template <int WINDOW>
  static __device__ __forceinline__ F fast_tap(F l2, F l1, F c, F r1, F r2, const WindowConst<F>& k)
  {
    if (WINDOW == 0) return c;
    else if (WINDOW == 3) return __fma_rn(-k.k1, __dadd_rn(l1, r1), __fma_rn(c, k.k0, __dadd_rn(l2, r2)));
    else return __fma_rn(c, k.k0, -__dadd_rn(l1, r1));
  }
  template <int WINDOW>
  static __device__ __forceinline__ cx<F> fast_window(cx<F> l2, cx<F> l1, cx<F> c, cx<F> r1, cx<F> r2, const WindowConst<F>& k)
  {
    cx<F> o;
    o.r = fast_tap<WINDOW>(l2.r, l1.r, c.r, r1.r, r2.r, k);
    o.i = fast_tap<WINDOW>(l2.i, l1.i, c.i, r1.i, r2.i, k);
    return o;
  }
};

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
/* one 32-byte group: 2 double bins or 4 float bins */
__device__ __forceinline__ void store_group(cx<double>* dst, const cx<double>* y)
{
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1, %2, %3, %4};"
               :: "l"(dst), "d"(y[0].r), "d"(y[0].i), "d"(y[1].r), "d"(y[1].i));
}
__device__ __forceinline__ void store_group(cx<float>* dst, const cx<float>* y)
{
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
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

/* accumulate / demodulate / window stages of the modulated replay: the reference's own roundings, or
 * (float MODE_FAST) the fused forms */
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
template <> struct StageOps<float, true>
{
  typedef float F;
  static __device__ __forceinline__ cx<F> mac(cx<F> acc, cx<F> p, F d) { return Arith<F>::mac_fused(acc, p, d); }
  static __device__ __forceinline__ cx<F> demod(cx<F> a, cx<F> p) { return Arith<F>::demod_fused(a, p); }
  template <int WINDOW>
  static __device__ __forceinline__ cx<F> window(cx<F> l2, cx<F> l1, cx<F> c, cx<F> r1, cx<F> r2, const WindowConst<F>& k)
  {
    return Arith<F>::template window_fused<WINDOW>(l2, l1, c, r1, r2, k);
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
  __device__ __forceinline__ int setup(unsigned group, unsigned lane, unsigned m)
  {
    const int e0 = (int)(group * G::SPAN) + 2 - G::HALO + (int)(lane * G::CPL);
#pragma unroll
    for (int b = 0; b < G::CPL; ++b)
    {
      const int slot = (int)lane * G::CPL + b;
      const int e = e0 + b;
      ok[b] = (slot >= G::HALO) && (slot < G::WC - G::HALO) && (e >= 2) && (e < (int)m + 2);
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

  /* one time step of the modulated replay: windowed spectrum of this lane's cells into y[] */
  template <bool RESTART, bool FUSED>
  __device__ __forceinline__ void compute(F d, const cx<F>* restart, const WindowConst<F>& win, cx<F>* y)
  {
    typedef Arith<F> A;
    typedef StageOps<F, FUSED> S;
    cx<F> x[G::CPL];
#pragma unroll
    for (int b = 0; b < G::CPL; ++b)
    {
      acc[b] = S::mac(acc[b], ph[b], d);
      ph[b] = RESTART ? restart[b] : A::rotate(ph[b], tw[b]);
      x[b] = S::demod(acc[b], ph[b]);
    }
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
 * Fused synthesis (EMIT_SYNTH): instead of storing the rows, every lane weighs its bins as sdft_isdft
 * does (sdft.h:639-652: (-1)^k Re(dft[k]) for latency 1, Re(dft[k] * tws[k]) otherwise) and the warp
 * reduces eight time steps at once with a transposing butterfly (9 shuffles per 8 steps instead of 5
 * per step): after three exchange rounds every lane holds ONE step's sum over eight lanes, two plain
 * butterfly rounds finish it.  The warp's partial sums over its bins go to part[group][t]; a second
 * tiny kernel adds the groups in order and scales by 2 (sdft.h:654-656).  Fixed order: deterministic.
 * ---------------------------------------------------------------------------------------------- */
template <typename F, int CPL, bool UNIT>
struct SynthLane
{
  F wr[CPL], wi[CPL];   // weights of this lane's bins; 0 for halo / out-of-range cells
  F p[8];

  __device__ __forceinline__ void setup(const cx<F>* __restrict__ tws, int e0, const bool* ok)
  {
#pragma unroll
    for (int b = 0; b < CPL; ++b)
    {
      const int k = e0 + b - 2;
      if (UNIT)
      {
        wr[b] = ok[b] ? ((k & 1) ? (F)(-1) : (F)(1)) : (F)0;
        wi[b] = (F)0;
      }
      else
      {
        wr[b] = ok[b] ? tws[k].r : (F)0;
        wi[b] = ok[b] ? tws[k].i : (F)0;
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) p[i] = (F)0;
  }

  __device__ __forceinline__ F weigh(const cx<F>* y) const
  {
    F s = (F)0;
#pragma unroll
    for (int b = 0; b < CPL; ++b)
    {
      s = fma(y[b].r, wr[b], s);
      if (!UNIT) s = fma(-y[b].i, wi[b], s);
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


/* ------------------------------------------------------------------------------------------------
 * K23  single-pass chained scan + emit (the production analysis kernel)
 *
 *      Work decomposition.  Time is cut into chunks (make_schedule), bins into warp-wide groups of
 *      Geo<F, GEO>::WC cells.  One WARP owns one (chunk, group); one CTA owns `W` CONSECUTIVE CHUNKS of the
 *      same group of one channel (a "block item").  Per warp:
 *        A. the chunk's own total  sum_i P[c+i] delta_i  (FP only, no memory traffic);
 *        B. the carry: totals of the CTA's chunks meet in shared memory; warp 0 adds them up in chunk
 *           order, publishes the CTA's aggregate, obtains the carry at the CTA's first chunk by a
 *           decoupled look-back over the PRECEDING CTAs of the same chain and publishes the inclusive
 *           prefix; every warp then adds the totals of the chunks before its own (shared memory again).
 *           The global chain is therefore W times shorter than the chunk chain, which is what bounds the
 *           latency of short calls (streaming, host tiles);
 *        C. replay the chunk from the carry, window across neighbouring cells, stream the rows out.
 *      Warps in phase A (FP only) and warps in phase C (store-bound) share every SM, so the scan
 *      arithmetic hides under the HBM-bound stores.
 *
 *      Ordering.  Every sum is taken in a fixed order that depends on the launch geometry only:
 *      carry(first chunk of CTA b) = ((acc + A_0) + A_1) + ... + A_{b-1} with A_c the CTA aggregates
 *      (each the in-order sum of its chunk totals), then + the totals of the CTA's earlier chunks in
 *      order.  The look-back walks back to the nearest CTA whose inclusive PREFIX is already published
 *      and adds the aggregates after it from left to right -- the same additions whatever that CTA
 *      happens to be, so results do not depend on timing.  Block items are handed out through an
 *      atomic ticket in (block, channel, group) order; an item only ever waits for items with smaller
 *      tickets, which have all started and publish their aggregates without waiting for anybody, so
 *      the kernel cannot deadlock whatever the block scheduler does.  Publication: cells are written
 *      by all lanes, fenced, then lane 0 releases a per-item flag stamped with the call's epoch (no
 *      flag clearing between calls); consumers acquire the flag and read the cells through L2.  A wait
 *      that exceeds kSpinLimitNs sets *error and gives up, so a logic error shows up as a reported
 *      failure, not as a hung device.
 * ---------------------------------------------------------------------------------------------- */
#ifndef SDFT_B200_EMIT_UNROLL
#define SDFT_B200_EMIT_UNROLL 2        // time steps unrolled in the row loop
#endif
#define SDFT_B200_STR2(x) #x
#define SDFT_B200_STR(x) SDFT_B200_STR2(x)
#define SDFT_B200_PRAGMA_UNROLL(n) _Pragma(SDFT_B200_STR(unroll n))
constexpr int kScanWarps = 8;          // most warps (= consecutive chunks) per scan/emit CTA
constexpr int kSmemSamples = 2048;     // deltas held per CTA: W * chunk length <= kSmemSamples
constexpr int kDeltaPad = 4;           // per-warp padding of the delta buffer: one zero sentinel, keeps 32-byte alignment

template <typename F> struct ChainArgs
{
  Schedule sched;
  const void* samples;     // (channels, sample_stride) time-domain samples of this call, float or double
  size_t sample_stride;
  const void* hist_old;    // (channels, 2m) the 2m samples before this call, oldest first
  void* hist_new;          // (channels, 2m) the 2m samples ending with this call's last one
  int td_double;           // time-domain type of samples/history: 0 float, 1 double
  F scale;                 // factor folded into the deltas (exactly 1 unless double MODE_FAST folds the window weight)
  const cx<F>* tw_ext;     // (cells)
  const cx<F>* f0;         // (rows, cells)
  const cx<F>* acc_in;     // (channels, cells)
  cx<F>* acc_out;
  cx<F>* totals;           // (channels, nblocks, groups, WC) aggregate of each block item
  cx<F>* prefix;           // (channels, nblocks, groups, WC) inclusive prefix after each block item
  unsigned* flags;         // (channels, nblocks, groups): 2*epoch = aggregate published, 2*epoch+1 = prefix published
  unsigned* control;       // [0] ticket counter, [1] error flag
  unsigned epoch;
  unsigned total_blocks;   // nblocks * channels * groups
  unsigned nblocks;        // block items per chain: ceil(nchunks / warps per CTA)
  unsigned channels;
  unsigned m;
  unsigned cells;
  cx<F>* out;              // (channels, n, m) or nullptr
  size_t out_channel_stride;
  const cx<F>* tws;        // (m) synthesis twiddles, EMIT_SYNTH only
  F* part;                 // (channels, groups, n) per-group partial sums of the fused synthesis
  unsigned groups;
  unsigned stage_rows;     // rows of look-back staging in shared memory (scan_stage_rows)
  WindowConst<F> win;
  unsigned long long* trace;   // -DSDFT_B200_TRACE builds only: 8 %globaltimer stamps per CTA, else unused
};

#if defined(SDFT_B200_TRACE)
#define SDFT_B200_STAMP(slot)                                                                          \
  do                                                                                                   \
  {                                                                                                    \
    if (a.trace && threadIdx.x == 0)                                                                   \
    {                                                                                                  \
      unsigned long long t__;                                                                          \
      asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t__));                                           \
      a.trace[(size_t)ticket * 8 + (slot)] = t__;                                                      \
    }                                                                                                  \
  } while (0)
#else
#define SDFT_B200_STAMP(slot) do { } while (0)
#endif

constexpr unsigned long long kSpinLimitNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void fence_acq_rel_gpu()
{
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
}
__device__ __forceinline__ void st_release_u32(unsigned* p, unsigned v)
{
  asm volatile("st.release.gpu.global.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
template <typename F> __device__ __forceinline__ cx<F> load_l2(const cx<F>* p);
template <> __device__ __forceinline__ cx<double> load_l2<double>(const cx<double>* p)
{
  const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
  cx<double> o; o.r = v.x; o.i = v.y; return o;
}
template <> __device__ __forceinline__ cx<float> load_l2<float>(const cx<float>* p)
{
  const float2 v = __ldcg(reinterpret_cast<const float2*>(p));
  cx<float> o; o.r = v.x; o.i = v.y; return o;
}
template <typename F> __device__ __forceinline__ void store_l2(cx<F>* p, cx<F> v);
template <> __device__ __forceinline__ void store_l2<double>(cx<double>* p, cx<double> v)
{
  __stcg(reinterpret_cast<double2*>(p), make_double2(v.r, v.i));
}
template <> __device__ __forceinline__ void store_l2<float>(cx<float>* p, cx<float> v)
{
  __stcg(reinterpret_cast<float2*>(p), make_float2(v.r, v.i));
}

enum { MODE_MODULATED = 0, MODE_FAST = 1 };
/* what phase C does with the windowed spectrum: nothing (state update only), store the (n, m) rows, or
 * feed the fused synthesis (latency == 1 / any latency, sdft.h:639) */
enum { EMIT_NONE = 0, EMIT_ROWS = 1, EMIT_SYNTH_UNIT = 2, EMIT_SYNTH = 3 };

/* float never runs the demodulated replay; these keep the shared kernel body compilable */
template <typename F, int MODE> struct FastOps
{
  static __device__ __forceinline__ cx<F> horner(cx<F> h, cx<F>, F) { return h; }
  static __device__ __forceinline__ cx<F> horner4(cx<F> h, cx<F>, cx<F>, cx<F>, cx<F>, F, F, F, F) { return h; }
  static __device__ __forceinline__ cx<F> cmul(cx<F> a, cx<F>) { return a; }
  static __device__ __forceinline__ cx<F> demod(cx<F> a, cx<F>) { return a; }
};
template <> struct FastOps<double, MODE_FAST>
{
  typedef double F;
  static __device__ __forceinline__ cx<F> horner(cx<F> h, cx<F> w, F d) { return Arith<F>::horner(h, w, d); }
  static __device__ __forceinline__ cx<F> horner4(cx<F> h, cx<F> w1, cx<F> w2, cx<F> w3, cx<F> w4, F d0, F d1, F d2, F d3)
  {
    return Arith<F>::horner4(h, w1, w2, w3, w4, d0, d1, d2, d3);
  }
  static __device__ __forceinline__ cx<F> cmul(cx<F> a, cx<F> b) { return Arith<F>::cmul(a, b); }
  static __device__ __forceinline__ cx<F> demod(cx<F> a, cx<F> p) { return Arith<F>::demod(a, p); }
};

/* K1 (fused prologue)  deltas of one chunk in TIME-DOMAIN precision (sdft.h:564, :186-191), loaded by
 *     the warp that owns the chunk:  ext(t) = history[t] for t < 2m, samples[t - 2m] otherwise;
 *     delta[t] = ext(t + 2m) - ext(t), one rounding in T, then widened to F. */
template <typename T, typename F>
__device__ __forceinline__ void chunk_deltas(const ChainArgs<F>& a, unsigned ch, const ChunkSpan& cs, F* sdelta, unsigned lane)
{
  const unsigned period = a.sched.period;
  const T* x = (const T*)a.samples + (size_t)ch * a.sample_stride;
  const T* ho = (const T*)a.hist_old + (size_t)ch * period;
  for (unsigned i = lane; i < cs.len; i += 32)
  {
    const unsigned long long t = cs.t0 + i;
    const T newest = x[t];
    const T oldest = (t < period) ? ho[t] : x[t - period];
    const T diff = newest - oldest;
    sdelta[i] = (F)diff * a.scale;
  }
}
/* the history the next call starts from; entries are dealt out over the block items of group 0 */
template <typename T, typename F>
__device__ __forceinline__ void roll_history(const ChainArgs<F>& a, unsigned ch, unsigned jb)
{
  const unsigned period = a.sched.period;
  const T* x = (const T*)a.samples + (size_t)ch * a.sample_stride;
  const T* ho = (const T*)a.hist_old + (size_t)ch * period;
  T* hn = (T*)a.hist_new + (size_t)ch * period;
  for (unsigned i = jb * blockDim.x + threadIdx.x; i < period; i += a.nblocks * blockDim.x)
  {
    const unsigned long long pos = a.sched.n + i;   // position inside history || samples
    hn[i] = (pos < period) ? ho[pos] : x[pos - period];
  }
}

/* true when the kernel variant <F, MODE> uses the demodulated double replay */
template <typename F, int MODE> struct IsSlide { enum { value = 0 }; };
template <> struct IsSlide<double, MODE_FAST> { enum { value = 1 }; };

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src)
{
  const unsigned dst = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

/* carry at the first chunk of block item `jb` (jb > 0): decoupled look-back over the preceding block
 * items of the chain, deterministic left-to-right summation (one warp; see the header comment).
 * The walk stops at the nearest item `q` with a published inclusive prefix -- or at item 0, whose
 * prefix is by definition acc_in + aggregate(0), so nobody waits for item 0's second publication.
 * The rows to add (prefix or acc_in, then the aggregates q+1 .. jb-1) are fetched into the shared-memory
 * staging area `stage` (`stage_rows` rows) with cp.async, as many at once as fit -- one memory round
 * trip for up to stage_rows rows instead of one per four -- and then added in order. */
template <typename F, int GEO>
__device__ __forceinline__ void look_back(const ChainArgs<F>& a, size_t item, size_t item_stride, unsigned jb, unsigned lane,
                                          const cx<F>* acc_in_cells, cx<F>* stage, unsigned stage_rows, cx<F>* acc,
                                          unsigned trace_slot)
{
  typedef Geo<F, GEO> G;
  typedef Arith<F> A;
  const unsigned code_total = a.epoch * 2u, code_prefix = a.epoch * 2u + 1u;
  long long top = (long long)jb - 1;
  long long q = -1;
  bool from_start = false;          // summation starts from acc_in + aggregate(0)
  unsigned long long t_start = 0;
  unsigned spins = 0;
  while (true)
  {
    const long long idx = top - (long long)lane;
    unsigned f = 0;
    if (idx >= 0) f = ld_relaxed_u32(a.flags + (item - (size_t)(jb - idx) * item_stride));   // acquire fence after the loop
    const bool is_prefix = (idx >= 0) && (f == code_prefix);
    const bool is_none = (idx >= 0) && (f != code_prefix) && (f != code_total);
    const unsigned mask_prefix = __ballot_sync(0xffffffffu, is_prefix);
    const unsigned mask_none = __ballot_sync(0xffffffffu, is_none);
    const unsigned mask_valid = __ballot_sync(0xffffffffu, idx >= 0);
    if (mask_prefix)
    {
      const int first = __ffs(mask_prefix) - 1;
      if ((mask_none & ((1u << first) - 1u)) == 0u)
      {
        q = top - first;
        break;
      }
    }
    else if (mask_none == 0u)
    {
      if (mask_valid != 0xffffffffu)
      {
        /* the window reaches item 0 and everything in it has at least its aggregate */
        q = 0;
        from_start = true;
        break;
      }
      top -= 32;      // 32 aggregates and no prefix yet: look further back
      continue;
    }
    /* a predecessor in the window has published nothing yet: wait for it (spin first, it is usually
     * a matter of a microsecond; back off and watch the clock only when it takes longer) */
    if (++spins < 64u) continue;
    __nanosleep(100);
    if (t_start == 0) t_start = global_timer_ns();
    else if (global_timer_ns() - t_start > kSpinLimitNs)
    {
      if (lane == 0) atomicExch(&a.control[1], 1u);
      q = 0;
      from_start = true;
      break;
    }
  }
#if defined(SDFT_B200_TRACE)
  if (a.trace && lane == 0) a.trace[(size_t)trace_slot * 8 + 7] = global_timer_ns();   // predecessors' publications seen
#endif
  fence_acq_rel_gpu();      // pairs with the publishers' st.release: their rows are visible from here on
  const cx<F>* chain0 = a.totals + (item - (size_t)jb * item_stride) * G::WC + lane * G::CPL;   // aggregate of item 0, this lane's cells
  const cx<F>* prefix0 = a.prefix + (item - (size_t)jb * item_stride) * G::WC + lane * G::CPL;
  const size_t rstride = item_stride * G::WC;
  if (from_start)
  {
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) acc[b] = acc_in_cells[b];
  }
  else
  {
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) acc[b] = load_l2<F>(prefix0 + (size_t)q * rstride + b);
  }
  long long r = from_start ? 0 : q + 1;
  cx<F>* mine = stage + lane * G::CPL;
  constexpr int kVec = (int)(G::CPL * sizeof(cx<F>) / 16);      // 16-byte pieces of this lane's cells in one row
  while (r < (long long)jb)
  {
    const unsigned batch = (unsigned)min((long long)stage_rows, (long long)jb - r);
    for (unsigned u = 0; u < batch; ++u)
    {
      const char* src = reinterpret_cast<const char*>(chain0 + (size_t)(r + u) * rstride);
      char* dst = reinterpret_cast<char*>(mine + (size_t)u * G::WC);
#pragma unroll
      for (int v = 0; v < kVec; ++v) cp_async_16(dst + 16 * v, src + 16 * v);
    }
    cp_async_wait_all();
    for (unsigned u = 0; u < batch; ++u)
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) acc[b] = A::cadd(acc[b], mine[(size_t)u * G::WC + b]);
    r += batch;
  }
}

template <typename F, int WINDOW, bool VEC, int EMIT, int MODE, int GEO>
__global__ void __launch_bounds__(kScanWarps * 32, 2) scan_emit_kernel(const ChainArgs<F> a)
{
  typedef EmitGeo<F, WINDOW, GEO> G;
  typedef Arith<F> A;
  constexpr bool SLIDE = IsSlide<F, MODE>::value != 0;     // double fast mode
  constexpr bool FUSED = (MODE == MODE_FAST) && !SLIDE;     // float fused mode
  typedef StageOps<F, FUSED> S;
  /* dynamic shared memory, sized by the launch (scan_smem_bytes): deltas of the CTA's chunks, their
   * totals, the carry at the CTA's first chunk */
  extern __shared__ __align__(32) unsigned char smem_raw[];
  __shared__ unsigned s_ticket;
  const unsigned nwarps = blockDim.x >> 5;
  F* sdelta_all = reinterpret_cast<F*>(smem_raw);
  cx<F>* stot_all = reinterpret_cast<cx<F>*>(smem_raw + (size_t)nwarps * (a.sched.chunk + kDeltaPad) * sizeof(F));
  cx<F>* scarry = stot_all + (size_t)nwarps * G::WC;
  cx<F>* sstage = scarry + G::WC;                            // look-back staging, a.stage_rows rows
#define stot(u) (stot_all + (size_t)(u) * G::WC)

  /* programmatic dependent launch (see launch_chain): nothing of the previous kernel in the stream may be
   * read or overwritten before it has completed; dependents of THIS kernel may start filling SMs as
   * soon as every CTA of it has got this far */
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0)
  {
    const unsigned t = atomicAdd(&a.control[0], 1u);
    if (t == a.total_blocks - 1) a.control[0] = 0;   // last ticket of the launch: rearm for the next call
    s_ticket = t;
  }
  __syncthreads();
  const unsigned ticket = s_ticket;
  SDFT_B200_STAMP(0);   // ticket taken
  /* (block item, channel, group): the chains of all channels and groups advance together */
  const unsigned per_block = a.channels * a.groups;
  const unsigned jb = ticket / per_block;
  const unsigned rem = ticket - jb * per_block;
  const unsigned ch = rem / a.groups;
  const unsigned group = rem - ch * a.groups;

  const unsigned warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31;
  const unsigned j = jb * nwarps + warp;                   // this warp's chunk
  const bool valid = j < a.sched.nchunks;
  const unsigned nvalid = min(nwarps, a.sched.nchunks - jb * nwarps);   // chunks of this CTA
  const bool last_block = (jb == a.nblocks - 1);
  ChunkSpan cs = chunk_span(a.sched, valid ? j : 0);
  F* sdelta = sdelta_all + warp * (a.sched.chunk + kDeltaPad);   // [len] holds a zero sentinel, see fast_compute

  if (valid)
  {
    if (a.td_double) chunk_deltas<double, F>(a, ch, cs, sdelta, lane);
    else chunk_deltas<float, F>(a, ch, cs, sdelta, lane);
    if (lane < 2) sdelta[cs.len + lane] = (F)0;
  }
  if (group == 0)
  {
    if (a.td_double) roll_history<double, F>(a, ch, jb);
    else roll_history<float, F>(a, ch, jb);
  }
  __syncwarp();
  SDFT_B200_STAMP(1);   // deltas in shared memory

  EmitLane<F, WINDOW, VEC, GEO> L;
  const int e0 = L.setup(group, lane, a.m);
  bool live[G::CPL];
  cx<F> zero;
  zero.r = (F)0; zero.i = (F)0;
#pragma unroll
  for (int b = 0; b < G::CPL; ++b)
  {
    const int e = e0 + b;
    live[b] = (e >= 0) && (e < (int)a.cells);
    L.tw[b] = live[b] ? a.tw_ext[e] : zero;
  }

  /* ---- phase A: this chunk's total ---- */
  cx<F> tot[G::CPL];
#pragma unroll
  for (int b = 0; b < G::CPL; ++b) { tot[b] = zero; L.ph[b] = zero; }
  if (valid)
  {
    if constexpr (SLIDE)
    {
      /* total = P_start * sum_i tw^i delta_i, the inner sum by Horner from the chunk's last sample,
       * four samples per step once the remaining count is a multiple of four */
      typedef FastOps<F, MODE> X;
      /* the table row of the starting phase is fetched now so that its latency hides under the sum */
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
        L.ph[b] = live[b] ? a.f0[(size_t)(cs.cursor0 / kF0Stride) * a.cells + (e0 + b)] : zero;
      int i = (int)cs.len;
      for (int r = i & 3; r > 0; --r)
      {
        const F d = sdelta[--i];
#pragma unroll
        for (int b = 0; b < G::CPL; ++b) tot[b] = X::horner(tot[b], L.tw[b], d);
      }
      {
        cx<F> w2[G::CPL], w3[G::CPL], w4[G::CPL];
#pragma unroll
        for (int b = 0; b < G::CPL; ++b)
        {
          w2[b] = X::cmul(L.tw[b], L.tw[b]);
          w3[b] = X::cmul(w2[b], L.tw[b]);
          w4[b] = X::cmul(w2[b], w2[b]);
        }
        while (i > 0)
        {
          i -= 4;
          const F d0 = sdelta[i], d1 = sdelta[i + 1], d2 = sdelta[i + 2], d3 = sdelta[i + 3];
#pragma unroll
          for (int b = 0; b < G::CPL; ++b) tot[b] = X::horner4(tot[b], L.tw[b], w2[b], w3[b], w4[b], d0, d1, d2, d3);
        }
      }
      for (unsigned r = cs.cursor0 % kF0Stride; r > 0; --r)     // only the first chunk of a call starts off the table grid
#pragma unroll
        for (int b = 0; b < G::CPL; ++b) L.ph[b] = A::rotate(L.ph[b], L.tw[b]);
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) tot[b] = X::cmul(L.ph[b], tot[b]);
    }
    else
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
        L.ph[b] = live[b] ? phase_at<F>(a.f0, a.cells, e0 + b, cs.cursor0, L.tw[b]) : zero;
      const unsigned body = cs.len - 1;
#pragma unroll 2
      for (unsigned i = 0; i < body; ++i)
      {
        const F d = sdelta[i];
#pragma unroll
        for (int b = 0; b < G::CPL; ++b)
        {
          tot[b] = S::mac(tot[b], L.ph[b], d);
          L.ph[b] = A::rotate(L.ph[b], L.tw[b]);
        }
      }
      const F d = sdelta[body];
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) tot[b] = S::mac(tot[b], L.ph[b], d);
    }
  }

  SDFT_B200_STAMP(2);   // chunk total done
  /* ---- phase B: carries (see the header comment) ---- */
  const size_t item_stride = (size_t)a.channels * a.groups;          // distance between consecutive block items of a chain
  const size_t item = (size_t)jb * item_stride + (size_t)ch * a.groups + group;
  if (nwarps > 1)
  {
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) stot(warp)[lane * G::CPL + b] = tot[b];
    __syncthreads();
  }
  if (warp == 0)
  {
    /* aggregate of the CTA: its chunk totals added in chunk order */
    cx<F> agg[G::CPL];
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) agg[b] = tot[b];
    for (unsigned u = 1; u < nvalid; ++u)
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) agg[b] = A::cadd(agg[b], stot(u)[lane * G::CPL + b]);
    if (!last_block)
    {
      cx<F>* tp = a.totals + item * G::WC + lane * G::CPL;
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) store_l2<F>(tp + b, agg[b]);
      /* the warp barrier orders every lane's stores before lane 0's release store, and a release is
       * cumulative: whoever acquires the flag sees the whole row (one fence instead of 32) */
      __syncwarp();
      if (lane == 0) st_release_u32(a.flags + item, a.epoch * 2u);
    }
    SDFT_B200_STAMP(3);   // aggregate published
    cx<F> carry[G::CPL];
    {
      const cx<F>* ai = a.acc_in + (size_t)ch * a.cells;
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) carry[b] = live[b] ? ai[e0 + b] : zero;
    }
    if (jb > 0)
    {
      cx<F> start[G::CPL];
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) start[b] = carry[b];
      look_back<F, GEO>(a, item, item_stride, jb, lane, start, sstage, a.stage_rows, carry, ticket);
    }
    SDFT_B200_STAMP(4);   // carry known
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) agg[b] = A::cadd(carry[b], agg[b]);
    if (!last_block)
    {
      cx<F>* pp = a.prefix + item * G::WC + lane * G::CPL;
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) store_l2<F>(pp + b, agg[b]);
      __syncwarp();
      if (lane == 0) st_release_u32(a.flags + item, a.epoch * 2u + 1u);
    }
    else
    {
      /* accumulators the next call starts with (sdft.h:157) */
      cx<F>* ao = a.acc_out + (size_t)ch * a.cells;
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
        if (live[b]) ao[e0 + b] = agg[b];
    }
#pragma unroll
    for (int b = 0; b < G::CPL; ++b) L.acc[b] = carry[b];
    if (nwarps > 1)
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) scarry[lane * G::CPL + b] = carry[b];
    }
  }
  if (nwarps > 1)
  {
    __syncthreads();
    if (warp > 0)
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) L.acc[b] = scarry[lane * G::CPL + b];
      for (unsigned u = 0; u < warp; ++u)
#pragma unroll
        for (int b = 0; b < G::CPL; ++b) L.acc[b] = A::cadd(L.acc[b], stot(u)[lane * G::CPL + b]);
    }
  }
  SDFT_B200_STAMP(5);   // carries distributed, replay starts
  if (!valid) return;

  /* ---- phase C: replay from the carry and stream the rows out ---- */
  if (EMIT == EMIT_ROWS)
  {
    const size_t row_stride = a.m;
    L.dst = a.out + (size_t)ch * a.out_channel_stride + (size_t)cs.t0 * row_stride + ((long long)e0 - 2);
    if constexpr (SLIDE)
    {
      /* anchor the demodulated spectrum at the carry (L.ph still holds the chunk's starting phase),
       * then slide; the period's last step needs no special case: conj(tw)^(2m) = 1 */
      typedef FastOps<F, MODE> X;
      const F d_first = sdelta[0];
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        L.acc[b] = X::demod(L.acc[b], L.ph[b]);
        L.acc[b].r = __dadd_rn(L.acc[b].r, d_first);     // z_0 = aux_{-1} + delta_0, see fast_compute
        L.tw[b].i = -L.tw[b].i;
      }
#if defined(SDFT_B200_NO_PIPELINE)
SDFT_B200_PRAGMA_UNROLL(SDFT_B200_EMIT_UNROLL)
      for (unsigned i = 0; i < cs.len; ++i) L.fast_step(sdelta[i + 1], a.win, row_stride);
#else
      {
        const F d1 = sdelta[1];
#pragma unroll
        for (int b = 0; b < G::CPL; ++b) L.acc[b] = A::horner(L.acc[b], L.tw[b], d1);      // z_1
      }
SDFT_B200_PRAGMA_UNROLL(SDFT_B200_EMIT_UNROLL)
      for (unsigned i = 0; i < cs.len; ++i)
      {
        cx<F> y[G::CPL];
        L.fast_compute_ahead(sdelta[i + 1], sdelta[i + 2], a.win, y);   // [len], [len + 1] are zero sentinels
        L.store_rows(y, row_stride);
      }
#endif
    }
    else
    {
      /* the starting phase is generated again rather than kept in registers across phase A */
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
        L.ph[b] = live[b] ? phase_at<F>(a.f0, a.cells, e0 + b, cs.cursor0, L.tw[b]) : zero;
      const unsigned body = cs.wraps ? cs.len - 1 : cs.len;
SDFT_B200_PRAGMA_UNROLL(SDFT_B200_EMIT_UNROLL)
      for (unsigned i = 0; i < body; ++i)
      {
        L.template step<false, FUSED>(sdelta[i], (const cx<F>*)nullptr, a.win, row_stride);
      }
      if (cs.wraps)
      {
        cx<F> restart[G::CPL];
#pragma unroll
        for (int b = 0; b < G::CPL; ++b) restart[b] = live[b] ? a.f0[e0 + b] : zero;
        L.template step<true, FUSED>(sdelta[body], restart, a.win, row_stride);
      }
    }
  }
  else if (EMIT == EMIT_SYNTH_UNIT || EMIT == EMIT_SYNTH)
  {
    /* fused synthesis: the rows never leave the registers (see SynthLane) */
    typedef SynthLane<F, G::CPL, EMIT == EMIT_SYNTH_UNIT> Y;
    Y syn;
    syn.setup(a.tws, e0, L.ok);
    F* pdst = a.part + ((size_t)ch * a.groups + group) * a.sched.n + cs.t0;
    const unsigned slot = Y::step_of(lane);
    const bool writer = (lane & 3u) == 0u;
    cx<F> restart[G::CPL];
    if constexpr (SLIDE)
    {
      typedef FastOps<F, MODE> X;
      const F d_first = sdelta[0];
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        L.acc[b] = X::demod(L.acc[b], L.ph[b]);
        L.acc[b].r = L.acc[b].r + d_first;               // z_0 = aux_{-1} + delta_0, see fast_compute
        L.tw[b].i = -L.tw[b].i;
      }
    }
    else
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        L.ph[b] = live[b] ? phase_at<F>(a.f0, a.cells, e0 + b, cs.cursor0, L.tw[b]) : zero;
        restart[b] = live[b] ? a.f0[e0 + b] : zero;
      }
    }
    /* eight steps per reduction while they last, then single steps; no branch encloses a shuffle */
    const unsigned body = (SLIDE || !cs.wraps) ? cs.len : cs.len - 1;
    unsigned i = 0;
    for (; i + 8 <= body; i += 8)
    {
#pragma unroll
      for (unsigned u = 0; u < 8; ++u)
      {
        cx<F> y[G::CPL];
        if constexpr (SLIDE) L.fast_compute(sdelta[i + u + 1], a.win, y);
        else L.template compute<false, FUSED>(sdelta[i + u], restart, a.win, y);
        syn.p[u] = syn.weigh(y);
      }
      const F total = syn.reduce8(lane);
      if (writer) pdst[i + slot] = total;
    }
    for (; i < body; ++i)
    {
      cx<F> y[G::CPL];
      if constexpr (SLIDE) L.fast_compute(sdelta[i + 1], a.win, y);
      else L.template compute<false, FUSED>(sdelta[i], restart, a.win, y);
      const F total = Y::warp_sum(syn.weigh(y));
      if (lane == 0) pdst[i] = total;
    }
    if (!SLIDE && cs.wraps)
    {
      cx<F> y[G::CPL];
      L.template compute<true, FUSED>(sdelta[body], restart, a.win, y);
      const F total = Y::warp_sum(syn.weigh(y));
      if (lane == 0) pdst[body] = total;
    }
  }
  SDFT_B200_STAMP(6);   // warp 0 finished its rows
}
#undef stot

/* dynamic shared memory of one scan/emit CTA of `warps` warps and chunk length `chunk` */
template <typename F, int GEO>
inline size_t scan_smem_bytes(unsigned warps, unsigned chunk)
{
  return (size_t)warps * (chunk + kDeltaPad) * sizeof(F) + (size_t)(warps + 1) * Geo<F, GEO>::WC * sizeof(cx<F>);
}
/* rows of look-back staging that fit next to it under the 48 KiB a CTA gets without opting in */
template <typename F, int GEO>
inline unsigned scan_stage_rows(unsigned warps, unsigned chunk)
{
  const size_t row = Geo<F, GEO>::WC * sizeof(cx<F>);
  const size_t base = scan_smem_bytes<F, GEO>(warps, chunk);
  size_t rows = ((size_t)48 * 1024 - base) / row;
  if (rows > 16) rows = 16;
  if (rows < 2) rows = 2;
  return (unsigned)rows;
}

/* ------------------------------------------------------------------------------------------------
 * K4  synthesis (sdft.h:635-672): one warp per row, lanes stride over bins, shuffle reduction.
 *     latency == 1: y = 2 * sum_k Re(dft[k]) * (-1)^k ; otherwise y = 2 * sum_k Re(dft[k] * tws[k]).
 *     The reference adds bins sequentially; the warp adds them in a fixed tree order instead
 *     (deterministic, differs by rounding only).
 * ---------------------------------------------------------------------------------------------- */
template <typename F>
__device__ __forceinline__ cx<F> load_stream(const cx<F>* p);
template <>
__device__ __forceinline__ cx<double> load_stream<double>(const cx<double>* p)
{
  cx<double> v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0, %1}, [%2];" : "=d"(v.r), "=d"(v.i) : "l"(p));
  return v;
}
template <>
__device__ __forceinline__ cx<float> load_stream<float>(const cx<float>* p)
{
  cx<float> v;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f32 {%0, %1}, [%2];" : "=f"(v.r), "=f"(v.i) : "l"(p));
  return v;
}

constexpr int kSynthWarps = 8;

template <typename T, typename F, bool UNIT_LATENCY>
__global__ void __launch_bounds__(kSynthWarps * 32) synth_kernel(const cx<F>* __restrict__ dfts,
                                                                 size_t dft_channel_stride,
                                                                 const cx<F>* __restrict__ tws,
                                                                 T* __restrict__ samples, size_t sample_stride,
                                                                 unsigned long long n, unsigned m)
{
  const unsigned ch = blockIdx.y;
  const unsigned lane = threadIdx.x & 31;
  const unsigned long long warps = (unsigned long long)gridDim.x * kSynthWarps;
  const cx<F>* base = dfts + (size_t)ch * dft_channel_stride;
  T* y = samples + (size_t)ch * sample_stride;
  for (unsigned long long row = (unsigned long long)blockIdx.x * kSynthWarps + (threadIdx.x >> 5); row < n; row += warps)
  {
    const cx<F>* r = base + (size_t)row * m;
    F s0 = (F)0, s1 = (F)0, s2 = (F)0, s3 = (F)0;
    unsigned k = lane;
    for (; k + 96 < m; k += 128)
    {
      const cx<F> v0 = load_stream<F>(r + k);
      const cx<F> v1 = load_stream<F>(r + k + 32);
      const cx<F> v2 = load_stream<F>(r + k + 64);
      const cx<F> v3 = load_stream<F>(r + k + 96);
      if (UNIT_LATENCY)
      {
        s0 += v0.r; s1 += v1.r; s2 += v2.r; s3 += v3.r;
      }
      else
      {
        const cx<F> w0 = tws[k], w1 = tws[k + 32], w2 = tws[k + 64], w3 = tws[k + 96];
        s0 += v0.r * w0.r - v0.i * w0.i;
        s1 += v1.r * w1.r - v1.i * w1.i;
        s2 += v2.r * w2.r - v2.i * w2.i;
        s3 += v3.r * w3.r - v3.i * w3.i;
      }
    }
    for (; k < m; k += 32)
    {
      const cx<F> v = load_stream<F>(r + k);
      if (UNIT_LATENCY)
      {
        s0 += v.r;
      }
      else
      {
        const cx<F> w = tws[k];
        s0 += v.r * w.r - v.i * w.i;
      }
    }
    F s = (s0 + s1) + (s2 + s3);
    /* k = lane + 32 i has the parity of the lane: apply (-1)^k once per lane */
    if (UNIT_LATENCY && (lane & 1)) s = -s;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if (lane == 0) y[row] = (T)(s * (F)2);
  }
}

}  // namespace sdftb200
