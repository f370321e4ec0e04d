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
 * Arithmetic policy: float frequency-domain data reproduces the reference's
 * rounding points with un-fused _rn intrinsics; double uses explicit FMAs in a FIXED pattern so that
 * every kernel generates bit-identical phases (the telescoping of x[t] - x[t-2m] depends on it).
 */
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace sdftb200
{

constexpr int kF0Stride = 32;     // phase table holds P at every 32nd cursor
constexpr int kMaxChunk = 1024;   // longest chunk the kernels accept (samples)
constexpr int kAutoChunk = 512;   // longest chunk the heuristic picks (measured best on B200)
constexpr int kEmitWarps = 4;     // warps per emit CTA
constexpr int kCellsPerLane = 4;  // consecutive cells owned by one lane
constexpr int kWarpCells = 32 * kCellsPerLane;

template <typename F> struct cx { F r, i; };

/* ------------------------------------------------------------------------------------------------
 * arithmetic policies
 * ---------------------------------------------------------------------------------------------- */
template <typename F> struct Arith;

template <> struct Arith<float>
{
  typedef float F;
  static __device__ __forceinline__ F add(F a, F b) { return __fadd_rn(a, b); }
  static __device__ __forceinline__ F sub(F a, F b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ F mul(F a, F b) { return __fmul_rn(a, b); }
  /* P * tw, sdft.h:584 via :298-300 -- bit-exact with the reference */
  static __device__ __forceinline__ cx<F> rotate(cx<F> p, cx<F> w)
  {
    cx<F> o;
    o.r = __fsub_rn(__fmul_rn(p.r, w.r), __fmul_rn(p.i, w.i));
    o.i = __fadd_rn(__fmul_rn(p.r, w.i), __fmul_rn(p.i, w.r));
    return o;
  }
  /* acc + P * delta, sdft.h:583 */
  static __device__ __forceinline__ cx<F> mac(cx<F> acc, cx<F> p, F d)
  {
    cx<F> o;
    o.r = __fadd_rn(acc.r, __fmul_rn(p.r, d));
    o.i = __fadd_rn(acc.i, __fmul_rn(p.i, d));
    return o;
  }
  /* acc * conj(P), sdft.h:585 */
  static __device__ __forceinline__ cx<F> demod(cx<F> a, cx<F> p)
  {
    const F ci = -p.i;
    cx<F> o;
    o.r = __fsub_rn(__fmul_rn(a.r, p.r), __fmul_rn(a.i, ci));
    o.i = __fadd_rn(__fmul_rn(a.r, ci), __fmul_rn(a.i, p.r));
    return o;
  }
};

template <> struct Arith<double>
{
  typedef double F;
  static __device__ __forceinline__ F add(F a, F b) { return __dadd_rn(a, b); }
  static __device__ __forceinline__ F sub(F a, F b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ F mul(F a, F b) { return __dmul_rn(a, b); }
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
};

/* window taps on one component, sdft.h:350-402; WINDOW follows enum sdft_window */
template <typename F> struct WindowConst
{
  F w;       // analysis weight 1/(2m)                        sdft.h:422
  F wq;      // w * 0.25, the Hann factor                     sdft.h:371
  F c0, c1, c2;   // double path: weight folded into the centre / first / second neighbour coefficients
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
  return k;
}

/* float: the reference's operation order, un-fused */
template <int WINDOW>
__device__ __forceinline__ float window_tap(float l2, float l1, float c, float r1, float r2, const WindowConst<float>& k)
{
  typedef Arith<float> A;
  typedef float F;
  if (WINDOW == 1)
  {
    return A::mul(A::sub(A::add(c, c), A::add(l1, r1)), k.wq);
  }
  else if (WINDOW == 2)
  {
    return A::mul(A::sub(A::mul(c, (F)(0.54)), A::mul(A::add(l1, r1), (F)(0.23))), k.w);
  }
  else if (WINDOW == 3)
  {
    const F a = A::mul(c, (F)(0.42));
    const F b = A::mul(A::add(l1, r1), (F)(0.25));
    const F d = A::mul(A::add(l2, r2), (F)(0.04));
    return A::mul(A::add(A::sub(a, b), d), k.w);
  }
  else
  {
    return A::mul(c, k.w);
  }
}

/* double: same taps with the weight folded into the coefficients and FMAs (3 / 3 / 5 FP64 instructions
 * per component instead of 4 / 5 / 8); differs from the reference's order by rounding only (~1e-16) */
template <int WINDOW>
__device__ __forceinline__ double window_tap(double l2, double l1, double c, double r1, double r2,
                                             const WindowConst<double>& k)
{
  if (WINDOW == 0)
  {
    return __dmul_rn(c, k.c0);
  }
  else if (WINDOW == 3)
  {
    const double t = __fma_rn(c, k.c0, -__dmul_rn(__dadd_rn(l1, r1), k.c1));
    return __fma_rn(__dadd_rn(l2, r2), k.c2, t);
  }
  else
  {
    return __fma_rn(c, k.c0, -__dmul_rn(__dadd_rn(l1, r1), k.c1));
  }
}

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
  unsigned f0_row;        // row of the phase table holding P at the chunk's first cursor
  bool first;             // chunk 0 of the call: phase comes from the plan's saved phase
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
  c.f0_row = (r * s.chunk) / kF0Stride;
  c.first = (j == 0);
  c.wraps = (ue == pe);
  return c;
}

/* mirror cells: source bin (or -1 = always zero) and whether the copy is conjugated.  Resolved on
 * the host from the assignment order of sdft.h:589-595 (matters only for m < 3). */
struct MirrorMap
{
  int cell[4];
  int src[4];
  int conj[4];
};

template <typename F>
__device__ __forceinline__ void store_with_mirrors(cx<F>* row, unsigned k, cx<F> v, const MirrorMap& mm)
{
  row[k + 2] = v;
#pragma unroll
  for (int q = 0; q < 4; ++q)
  {
    if (mm.src[q] == (int)k)
    {
      cx<F> c = v;
      if (mm.conj[q]) c.i = -c.i;
      row[mm.cell[q]] = c;
    }
    else if (mm.src[q] < 0 && k == 0)
    {
      cx<F> z;
      z.r = (F)0; z.i = (F)0;
      row[mm.cell[q]] = z;
    }
  }
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

/* ------------------------------------------------------------------------------------------------
 * K1  deltas in TIME-DOMAIN precision (sdft.h:564) and the new 2m-sample history
 *     ext(t) = history[t] for t < 2m, samples[t - 2m] otherwise; delta[t] = ext(t + 2m) - ext(t)
 * ---------------------------------------------------------------------------------------------- */
template <typename T, typename F>
__global__ void delta_kernel(const T* __restrict__ samples, size_t sample_stride,
                             const T* __restrict__ hist_old, T* __restrict__ hist_new,
                             F* __restrict__ delta, size_t delta_stride,
                             unsigned long long n, unsigned period)
{
  const unsigned ch = blockIdx.y;
  const T* x = samples + (size_t)ch * sample_stride;
  const T* ho = hist_old + (size_t)ch * period;
  T* hn = hist_new + (size_t)ch * period;
  F* d = delta + (size_t)ch * delta_stride;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += stride)
  {
    const T newest = x[t];
    const T oldest = (t < period) ? ho[t] : x[t - period];
    const T diff = newest - oldest;   // T is float or double: one rounding in TD precision
    d[t] = (F)diff;
  }
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < period; i += stride)
  {
    const unsigned long long pos = n + i;   // position inside history ‖ samples
    hn[i] = (pos < period) ? ho[pos] : x[pos - period];
  }
}

/* ------------------------------------------------------------------------------------------------
 * K2  chunk totals (scan pass 1): S[ch][j][k+2] = sum_i P[c_j + i][k] * delta[t_j + i]
 *     one thread per bin; the last chunk also leaves the plan's phase for the next call
 * ---------------------------------------------------------------------------------------------- */
template <typename F> struct ScanArgs
{
  Schedule sched;
  const F* delta;          // (channels, delta_stride)
  size_t delta_stride;
  const cx<F>* tw_ext;     // (cells)
  const cx<F>* f0;         // (rows, cells)
  const cx<F>* phase_in;   // (channels, cells)  P at the cursor the call starts with
  cx<F>* phase_out;        // (channels, cells)  P at the cursor the call ends with (other buffer)
  cx<F>* acc_state;        // (channels, cells)
  cx<F>* totals;           // (channels, nchunks, cells): totals, then carries in place
  unsigned m;
  unsigned cells;          // m + 4
  MirrorMap mirrors;
};

constexpr int kTotalsThreads = 128;

template <typename F>
__global__ void __launch_bounds__(kTotalsThreads) chunk_totals_kernel(const ScanArgs<F> a)
{
  __shared__ F sdelta[kMaxChunk];
  const unsigned bin_blocks = (a.m + kTotalsThreads - 1) / kTotalsThreads;
  const unsigned j = blockIdx.x / bin_blocks;
  const unsigned bb = blockIdx.x - j * bin_blocks;
  const unsigned ch = blockIdx.y;
  const ChunkSpan cs = chunk_span(a.sched, j);

  const F* dsrc = a.delta + (size_t)ch * a.delta_stride + cs.t0;
  for (unsigned i = threadIdx.x; i < cs.len; i += kTotalsThreads) sdelta[i] = dsrc[i];
  __syncthreads();

  const unsigned k = bb * kTotalsThreads + threadIdx.x;
  if (k >= a.m) return;
  const unsigned e = k + 2;
  const cx<F> w = a.tw_ext[e];
  cx<F> p = cs.first ? a.phase_in[(size_t)ch * a.cells + e] : a.f0[(size_t)cs.f0_row * a.cells + e];
  cx<F> acc;
  acc.r = (F)0; acc.i = (F)0;
  const unsigned body = cs.len - 1;
#pragma unroll 4
  for (unsigned i = 0; i < body; ++i)
  {
    acc = Arith<F>::mac(acc, p, sdelta[i]);
    p = Arith<F>::rotate(p, w);
  }
  acc = Arith<F>::mac(acc, p, sdelta[body]);
  a.totals[((size_t)ch * a.sched.nchunks + j) * a.cells + e] = acc;

  if (j == a.sched.nchunks - 1)
  {
    p = cs.wraps ? a.f0[e] : Arith<F>::rotate(p, w);   // row 0 of the table is the restart value
    store_with_mirrors(a.phase_out + (size_t)ch * a.cells, k, p, a.mirrors);
  }
}

/* ------------------------------------------------------------------------------------------------
 * K2b  exclusive scan of the chunk totals over time, seeded with the plan's accumulators
 *      (the running accoutput of sdft.h:157); totals[] becomes the carry entering each chunk
 * ---------------------------------------------------------------------------------------------- */
template <typename F>
__global__ void carry_scan_kernel(const ScanArgs<F> a)
{
  const unsigned k = blockIdx.x * blockDim.x + threadIdx.x;
  const unsigned ch = blockIdx.y;
  if (k >= a.m) return;
  cx<F>* acc_row = a.acc_state + (size_t)ch * a.cells;
  cx<F> run = acc_row[k + 2];
  cx<F>* base = a.totals + (size_t)ch * a.sched.nchunks * a.cells;
  const unsigned n = a.sched.nchunks;
  constexpr int U = 8;
  unsigned j = 0;
  for (; j + U <= n; j += U)
  {
    cx<F> v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) v[u] = base[(size_t)(j + u) * a.cells + k + 2];
#pragma unroll
    for (int u = 0; u < U; ++u)
    {
      store_with_mirrors(base + (size_t)(j + u) * a.cells, k, run, a.mirrors);
      run.r = Arith<F>::add(run.r, v[u].r);
      run.i = Arith<F>::add(run.i, v[u].i);
    }
  }
  for (; j < n; ++j)
  {
    const cx<F> v = base[(size_t)j * a.cells + k + 2];
    store_with_mirrors(base + (size_t)j * a.cells, k, run, a.mirrors);
    run.r = Arith<F>::add(run.r, v.r);
    run.i = Arith<F>::add(run.i, v.i);
  }
  store_with_mirrors(acc_row, k, run, a.mirrors);
}

/* ------------------------------------------------------------------------------------------------
 * K3  emit (scan pass 2, the dominant kernel): replay each chunk from its carry, demodulate, apply
 *     the window across neighbouring cells and stream the (n, m) rows out.
 *
 *     One warp owns kWarpCells = 128 consecutive cells (4 per lane) of one chunk and is independent
 *     of every other warp: the 2 outermost cells on either side are halo (recomputed by the
 *     neighbouring warp), so 124 bins per warp are stored (128 for the boxcar window).  Neighbour
 *     cells inside the warp come from registers or one shuffle.  Rows are written with consecutive
 *     lanes on consecutive bins; when m is even each lane stores aligned pairs of bins
 *     (32 B for double, 16 B for float) with an evict-first policy.
 * ---------------------------------------------------------------------------------------------- */
template <typename F> struct EmitArgs
{
  ScanArgs<F> scan;
  cx<F>* out;              // (channels, n, m)
  size_t out_channel_stride;   // in complex elements
  unsigned groups;         // warps needed to cover one row
  unsigned group_blocks;   // CTAs per chunk
  WindowConst<F> win;
};

__device__ __forceinline__ void store_pair(cx<double>* dst, cx<double> a, cx<double> b)
{
  asm volatile("st.global.L1::no_allocate.L2::evict_first.v4.f64 [%0], {%1, %2, %3, %4};"
               :: "l"(dst), "d"(a.r), "d"(a.i), "d"(b.r), "d"(b.i));
}
__device__ __forceinline__ void store_pair(cx<float>* dst, cx<float> a, cx<float> b)
{
  asm volatile("st.global.cs.v4.f32 [%0], {%1, %2, %3, %4};"
               :: "l"(dst), "f"(a.r), "f"(a.i), "f"(b.r), "f"(b.i));
}
__device__ __forceinline__ void store_one(cx<double>* dst, cx<double> a)
{
  asm volatile("st.global.cs.v2.f64 [%0], {%1, %2};"
               :: "l"(dst), "d"(a.r), "d"(a.i));
}
__device__ __forceinline__ void store_one(cx<float>* dst, cx<float> a)
{
  asm volatile("st.global.cs.v2.f32 [%0], {%1, %2};"
               :: "l"(dst), "f"(a.r), "f"(a.i));
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

template <typename F, int WINDOW, bool VEC>
struct EmitLane
{
  cx<F> acc[kCellsPerLane];
  cx<F> ph[kCellsPerLane];
  cx<F> tw[kCellsPerLane];
  cx<F>* dst;            // address of this lane's cell 0 in the current row (may be out of range)
  bool ok[kCellsPerLane];

  /* one time step; RESTART = the period's last step, after which the phase restarts (sdft.h:566-576) */
  template <bool RESTART>
  __device__ __forceinline__ void step(F d, const cx<F>* restart, const WindowConst<F>& win, size_t row_stride)
  {
    typedef Arith<F> A;
    cx<F> x[kCellsPerLane];
#pragma unroll
    for (int b = 0; b < kCellsPerLane; ++b)
    {
      acc[b] = A::mac(acc[b], ph[b], d);
      ph[b] = RESTART ? restart[b] : A::rotate(ph[b], tw[b]);
      x[b] = A::demod(acc[b], ph[b]);
    }
    cx<F> y[kCellsPerLane];
    if (WINDOW == 0)
    {
#pragma unroll
      for (int b = 0; b < kCellsPerLane; ++b)
      {
        y[b].r = window_tap<0>(x[b].r, x[b].r, x[b].r, x[b].r, x[b].r, win);
        y[b].i = window_tap<0>(x[b].i, x[b].i, x[b].i, x[b].i, x[b].i, win);
      }
    }
    else
    {
      /* neighbours: [l2 l1 | x0 x1 x2 x3 | r1 r2] */
      cx<F> l1 = shfl_up1(x[3]);
      cx<F> r1 = shfl_down1(x[0]);
      cx<F> l2, r2;
      if (WINDOW == 3)
      {
        l2 = shfl_up1(x[2]);
        r2 = shfl_down1(x[1]);
      }
      else
      {
        l2 = l1; r2 = r1;   // unused
      }
      y[0].r = window_tap<WINDOW>(l2.r, l1.r, x[0].r, x[1].r, x[2].r, win);
      y[0].i = window_tap<WINDOW>(l2.i, l1.i, x[0].i, x[1].i, x[2].i, win);
      y[1].r = window_tap<WINDOW>(l1.r, x[0].r, x[1].r, x[2].r, x[3].r, win);
      y[1].i = window_tap<WINDOW>(l1.i, x[0].i, x[1].i, x[2].i, x[3].i, win);
      y[2].r = window_tap<WINDOW>(x[0].r, x[1].r, x[2].r, x[3].r, r1.r, win);
      y[2].i = window_tap<WINDOW>(x[0].i, x[1].i, x[2].i, x[3].i, r1.i, win);
      y[3].r = window_tap<WINDOW>(x[1].r, x[2].r, x[3].r, r1.r, r2.r, win);
      y[3].i = window_tap<WINDOW>(x[1].i, x[2].i, x[3].i, r1.i, r2.i, win);
    }
    if (VEC)
    {
      if (ok[0]) store_pair(dst, y[0], y[1]);
      if (ok[2]) store_pair(dst + 2, y[2], y[3]);
    }
    else
    {
#pragma unroll
      for (int b = 0; b < kCellsPerLane; ++b)
        if (ok[b]) store_one(dst + b, y[b]);
    }
    dst += row_stride;
  }
};

template <typename F, int WINDOW, bool VEC>
__global__ void __launch_bounds__(kEmitWarps * 32) emit_kernel(const EmitArgs<F> a)
{
  constexpr int HALO = (WINDOW == 0) ? 0 : 2;
  constexpr int SPAN = kWarpCells - 2 * HALO;   // bins stored per warp
  __shared__ F sdelta[kMaxChunk];

  const ScanArgs<F>& s = a.scan;
  const unsigned j = blockIdx.x / a.group_blocks;
  const unsigned gblk = blockIdx.x - j * a.group_blocks;
  const unsigned ch = blockIdx.y;
  const ChunkSpan cs = chunk_span(s.sched, j);

  const F* dsrc = s.delta + (size_t)ch * s.delta_stride + cs.t0;
  for (unsigned i = threadIdx.x; i < cs.len; i += kEmitWarps * 32) sdelta[i] = dsrc[i];
  __syncthreads();

  const unsigned warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31;
  const unsigned group = gblk * kEmitWarps + warp;
  if (group >= a.groups) return;

  /* cell index of this lane's slot 0: the warp's 128 cells start HALO cells below its first bin */
  const unsigned first_bin = group * SPAN;
  const unsigned e0 = first_bin + 2 - HALO + lane * kCellsPerLane;

  EmitLane<F, WINDOW, VEC> L;
  cx<F> restart[kCellsPerLane];
  const cx<F>* phase_src = cs.first ? s.phase_in + (size_t)ch * s.cells
                                    : s.f0 + (size_t)cs.f0_row * s.cells;
  const cx<F>* carry_src = s.totals + ((size_t)ch * s.sched.nchunks + j) * s.cells;
#pragma unroll
  for (int b = 0; b < kCellsPerLane; ++b)
  {
    const unsigned e = e0 + b;
    const bool live = e < s.cells;
    cx<F> z;
    z.r = (F)0; z.i = (F)0;
    L.tw[b] = live ? s.tw_ext[e] : z;
    L.ph[b] = live ? phase_src[e] : z;
    L.acc[b] = live ? carry_src[e] : z;
    restart[b] = live ? s.f0[e] : z;
    /* stored iff the cell is one of the warp's SPAN inner cells and a real bin below m */
    const unsigned slot = lane * kCellsPerLane + b;
    const long long k = (long long)e - 2;
    L.ok[b] = ((int)slot >= HALO) && (slot < (unsigned)(kWarpCells - HALO)) && (e >= 2u) && (k < (long long)s.m);
  }
  const size_t row_stride = s.m;
  L.dst = a.out + (size_t)ch * a.out_channel_stride + (size_t)cs.t0 * row_stride + ((long long)e0 - 2);

  const unsigned body = cs.wraps ? cs.len - 1 : cs.len;
#pragma unroll 2
  for (unsigned i = 0; i < body; ++i)
  {
    L.template step<false>(sdelta[i], restart, a.win, row_stride);
  }
  if (cs.wraps)
  {
    L.template step<true>(sdelta[body], restart, a.win, row_stride);
  }
}

/* ------------------------------------------------------------------------------------------------
 * K23  single-pass chained scan + emit (the production analysis kernel)
 *
 *      Same work items as K3 (one warp = one chunk x 128 cells), but the warp first computes its own
 *      chunk total (K2's job, FP64/FP32 only, no memory traffic), then obtains the carry from the warp
 *      that owns the PREVIOUS chunk of the same cells, publishes its inclusive prefix, and only then
 *      replays the chunk and streams the rows out.  Warps in the compute phase and warps in the store
 *      phase share every SM, so the scan arithmetic hides under the HBM-bound stores instead of
 *      running as separate kernels in front of them.
 *
 *      Ordering.  carry_j = (((acc + total_0) + total_1) + ...) + total_{j-1}, always added in chunk
 *      order, so results are deterministic and independent of timing (see the look-back comment in the
 *      kernel).  Work items are handed out through an atomic ticket in (channel, chunk, group) order;
 *      an item only ever waits for items with smaller tickets, which have all started and publish
 *      their totals without waiting for anybody, so the kernel cannot deadlock whatever the block
 *      scheduler does.  Publication: cells are written by all lanes, fenced, then lane 0 releases a
 *      per-item flag stamped with the call's epoch (no flag clearing between calls); consumers acquire
 *      the flag and read the cells through L2.  A wait that exceeds kSpinLimitNs sets *error and gives
 *      up, so a logic error shows up as a reported failure, not as a hung device.
 * ---------------------------------------------------------------------------------------------- */
template <typename F> struct ChainArgs
{
  Schedule sched;
  const F* delta;          // (channels, delta_stride)
  size_t delta_stride;
  const cx<F>* tw_ext;     // (cells)
  const cx<F>* f0;         // (rows, cells)
  const cx<F>* phase_in;   // (channels, cells)
  cx<F>* phase_out;
  const cx<F>* acc_in;     // (channels, cells)
  cx<F>* acc_out;
  cx<F>* totals;           // (channels, nchunks, groups, kWarpCells) each chunk's own total
  cx<F>* prefix;           // (channels, nchunks, groups, kWarpCells) inclusive prefix after each chunk
  unsigned* flags;         // (channels, nchunks, groups): 2*epoch = total published, 2*epoch+1 = prefix published
  unsigned* control;       // [0] ticket counter, [1] error flag
  unsigned epoch;
  unsigned total_blocks;
  unsigned m;
  unsigned cells;
  cx<F>* out;              // (channels, n, m) or nullptr
  size_t out_channel_stride;
  unsigned groups;
  unsigned group_blocks;
  WindowConst<F> win;
};

constexpr unsigned long long kSpinLimitNs = 20ull * 1000ull * 1000ull * 1000ull;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p)
{
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
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

template <typename F, int WINDOW, bool VEC, bool EMIT>
__global__ void __launch_bounds__(kEmitWarps * 32, 4) scan_emit_kernel(const ChainArgs<F> a)
{
  constexpr int HALO = (WINDOW == 0) ? 0 : 2;
  constexpr int SPAN = kWarpCells - 2 * HALO;
  __shared__ F sdelta[kMaxChunk];
  __shared__ unsigned s_ticket;

  if (threadIdx.x == 0)
  {
    const unsigned t = atomicAdd(&a.control[0], 1u);
    if (t == a.total_blocks - 1) a.control[0] = 0;   // last ticket of the launch: rearm for the next call
    s_ticket = t;
  }
  __syncthreads();
  const unsigned ticket = s_ticket;
  const unsigned per_channel = a.sched.nchunks * a.group_blocks;
  const unsigned ch = ticket / per_channel;
  const unsigned rem = ticket - ch * per_channel;
  const unsigned j = rem / a.group_blocks;
  const unsigned gblk = rem - j * a.group_blocks;
  const ChunkSpan cs = chunk_span(a.sched, j);

  const F* dsrc = a.delta + (size_t)ch * a.delta_stride + cs.t0;
  for (unsigned i = threadIdx.x; i < cs.len; i += kEmitWarps * 32) sdelta[i] = dsrc[i];
  __syncthreads();

  const unsigned warp = threadIdx.x >> 5;
  const unsigned lane = threadIdx.x & 31;
  const unsigned group = gblk * kEmitWarps + warp;
  if (group >= a.groups) return;

  const unsigned first_bin = group * SPAN;
  const unsigned e0 = first_bin + 2 - HALO + lane * kCellsPerLane;
  const bool last_chunk = (j == a.sched.nchunks - 1);

  EmitLane<F, WINDOW, VEC> L;
  cx<F> restart[kCellsPerLane];
  cx<F> start[kCellsPerLane];
  bool live[kCellsPerLane];
  const cx<F>* phase_src = cs.first ? a.phase_in + (size_t)ch * a.cells : a.f0 + (size_t)cs.f0_row * a.cells;
  cx<F> zero;
  zero.r = (F)0; zero.i = (F)0;
#pragma unroll
  for (int b = 0; b < kCellsPerLane; ++b)
  {
    const unsigned e = e0 + b;
    live[b] = e < a.cells;
    L.tw[b] = live[b] ? a.tw_ext[e] : zero;
    start[b] = live[b] ? phase_src[e] : zero;
    restart[b] = live[b] ? a.f0[e] : zero;
    const unsigned slot = lane * kCellsPerLane + b;
    L.ok[b] = ((int)slot >= HALO) && (slot < (unsigned)(kWarpCells - HALO)) && (e >= 2u) && (e < a.m + 2u);
  }

  /* ---- phase A: this chunk's total ---- */
  cx<F> tot[kCellsPerLane];
#pragma unroll
  for (int b = 0; b < kCellsPerLane; ++b)
  {
    tot[b] = zero;
    L.ph[b] = start[b];
  }
  {
    const unsigned body = cs.len - 1;
#pragma unroll 2
    for (unsigned i = 0; i < body; ++i)
    {
      const F d = sdelta[i];
#pragma unroll
      for (int b = 0; b < kCellsPerLane; ++b)
      {
        tot[b] = Arith<F>::mac(tot[b], L.ph[b], d);
        L.ph[b] = Arith<F>::rotate(L.ph[b], L.tw[b]);
      }
    }
    const F d = sdelta[body];
#pragma unroll
    for (int b = 0; b < kCellsPerLane; ++b) tot[b] = Arith<F>::mac(tot[b], L.ph[b], d);
  }
  if (last_chunk)
  {
    /* phase the next call starts with (sdft.h:573 / :584) */
    cx<F>* po = a.phase_out + (size_t)ch * a.cells;
#pragma unroll
    for (int b = 0; b < kCellsPerLane; ++b)
      if (live[b]) po[e0 + b] = cs.wraps ? restart[b] : Arith<F>::rotate(L.ph[b], L.tw[b]);
  }

  /* ---- carry: decoupled look-back with a deterministic, left-to-right summation ----
   * Every item first publishes its own total ("aggregate"), which depends on nothing.  To obtain its
   * carry an item walks back over its predecessors' flags, 32 at a time, to the nearest one whose
   * inclusive PREFIX is already known, then adds prefix[q] + total[q+1] + ... + total[j-1] from left
   * to right.  That is the very sequence of additions the serial chain would perform, so the result
   * is bit-identical whatever q happens to be, but no item ever waits for a chain of predecessors. */
  const size_t item = ((size_t)ch * a.sched.nchunks + j) * a.groups + group;
  const size_t item_stride = a.groups;                      // distance between consecutive chunks
  const unsigned code_total = a.epoch * 2u, code_prefix = a.epoch * 2u + 1u;
  if (j == 0)
  {
    const cx<F>* ai = a.acc_in + (size_t)ch * a.cells;
#pragma unroll
    for (int b = 0; b < kCellsPerLane; ++b) L.acc[b] = live[b] ? ai[e0 + b] : zero;
  }
  else
  {
    if (!last_chunk)
    {
      cx<F>* tp = a.totals + item * kWarpCells + lane * kCellsPerLane;
#pragma unroll
      for (int b = 0; b < kCellsPerLane; ++b) store_l2<F>(tp + b, tot[b]);
      __threadfence();
      __syncwarp();
      if (lane == 0) st_release_u32(a.flags + item, code_total);
    }
    /* find q = nearest predecessor with a published prefix; all items in (q, j) must have totals */
    long long top = (long long)j - 1;
    long long q = -1;
    unsigned long long t_start = 0;
    while (true)
    {
      const long long idx = top - (long long)lane;
      unsigned f = 0;
      if (idx >= 0) f = ld_acquire_u32(a.flags + (item - (size_t)(j - idx) * item_stride));
      const bool is_prefix = (idx >= 0) && (f == code_prefix);
      const bool is_none = (idx >= 0) && (f != code_prefix) && (f != code_total);
      const unsigned mask_prefix = __ballot_sync(0xffffffffu, is_prefix);
      const unsigned mask_none = __ballot_sync(0xffffffffu, is_none);
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
        top -= 32;      // 32 totals and no prefix yet: look further back
        continue;
      }
      /* a predecessor in the window has published nothing yet: wait for it */
      __nanosleep(40);
      if (t_start == 0) t_start = global_timer_ns();
      else if (global_timer_ns() - t_start > kSpinLimitNs)
      {
        if (lane == 0) atomicExch(&a.control[1], 1u);
        q = 0;
        break;
      }
    }
    __threadfence();
    {
      const size_t qi = item - (size_t)(j - q) * item_stride;
      const cx<F>* pp = a.prefix + qi * kWarpCells + lane * kCellsPerLane;
#pragma unroll
      for (int b = 0; b < kCellsPerLane; ++b) L.acc[b] = load_l2<F>(pp + b);
      for (long long r = q + 1; r < (long long)j; ++r)
      {
        const size_t ri = item - (size_t)(j - r) * item_stride;
        const cx<F>* tp = a.totals + ri * kWarpCells + lane * kCellsPerLane;
        cx<F> v[kCellsPerLane];
#pragma unroll
        for (int b = 0; b < kCellsPerLane; ++b) v[b] = load_l2<F>(tp + b);
#pragma unroll
        for (int b = 0; b < kCellsPerLane; ++b)
        {
          L.acc[b].r = Arith<F>::add(L.acc[b].r, v[b].r);
          L.acc[b].i = Arith<F>::add(L.acc[b].i, v[b].i);
        }
      }
    }
  }
#pragma unroll
  for (int b = 0; b < kCellsPerLane; ++b)
  {
    tot[b].r = Arith<F>::add(L.acc[b].r, tot[b].r);
    tot[b].i = Arith<F>::add(L.acc[b].i, tot[b].i);
  }
  if (!last_chunk)
  {
    cx<F>* pp = a.prefix + item * kWarpCells + lane * kCellsPerLane;
#pragma unroll
    for (int b = 0; b < kCellsPerLane; ++b) store_l2<F>(pp + b, tot[b]);
    __threadfence();
    __syncwarp();
    if (lane == 0) st_release_u32(a.flags + item, code_prefix);
  }
  else
  {
    /* accumulators the next call starts with (sdft.h:157) */
    cx<F>* ao = a.acc_out + (size_t)ch * a.cells;
#pragma unroll
    for (int b = 0; b < kCellsPerLane; ++b)
      if (live[b]) ao[e0 + b] = tot[b];
  }

  /* ---- phase B: replay from the carry and stream the rows out ---- */
  if (EMIT)
  {
#pragma unroll
    for (int b = 0; b < kCellsPerLane; ++b) L.ph[b] = start[b];
    const size_t row_stride = a.m;
    L.dst = a.out + (size_t)ch * a.out_channel_stride + (size_t)cs.t0 * row_stride + ((long long)e0 - 2);
    const unsigned body = cs.wraps ? cs.len - 1 : cs.len;
#pragma unroll 2
    for (unsigned i = 0; i < body; ++i)
    {
      L.template step<false>(sdelta[i], restart, a.win, row_stride);
    }
    if (cs.wraps)
    {
      L.template step<true>(sdelta[body], restart, a.win, row_stride);
    }
  }
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
