/*
 * sdft_arith.cuh -- arithmetic policies: window constants, float (strict / fused) and double (modulated / fast) complex operations (c/src/sdft/sdft.h:298-300, :350-402).
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include "sdft_common.cuh"

namespace sdftb200
{

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

template <> struct Arith<float>
{
  typedef float F;
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


}  // namespace sdftb200
