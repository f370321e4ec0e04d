/*
 * sdft_scan.cuh -- K23, the single-pass chained scan + emit kernel that executes one sdft_sdft_n call (c/src/sdft/sdft.h:562-613).
 * Part of the sm_100a kernels of libsdft_b200.so; see sdft_kernels.cuh for the overview.
 */
#pragma once

#include "sdft_schedule.cuh"
#include "sdft_lane.cuh"

namespace sdftb200
{

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
 *
 *      Between calls.  A call hands three things to the next call on the plan: the 2m-sample history
 *      (written by the call's CTAs in their prologue, a slice each), the accumulators (written by the last block item
 *      of every chain) and, implicitly, the order of completion.  Each hand-over has its own counter
 *      (ChainArgs::sync: [0] history pieces written, [1] accumulator rows written; monotonic, the host knows
 *      the totals), so the consumer can wait for exactly what it needs:
 *        - serial call (flow == 0, the default): griddepcontrol.wait at the top -- the previous kernel has
 *          completed and flushed, every counter is trivially there, nothing is polled;
 *        - STREAMING call (flow == 1, sdft_b200_set_streaming): no wait at the top.  The kernel is launched
 *          with programmatic stream serialization, so its CTAs become resident as soon as every CTA of the
 *          previous call has started, compute their deltas and chunk totals while that call is still
 *          streaming rows out, and poll the counters only where the data is needed: the history before the
 *          deltas of the first 2m samples, the accumulators at the head of each chain.  Scratch (ticket,
 *          flags, totals, prefixes) and state buffers rotate through rings sized by the streaming depth D;
 *          the plan counts completed calls ([completed], bumped by the last CTA of every call), and a
 *          streaming call first makes sure that the call D calls back -- the previous user of its slot --
 *          is among them, which bounds the calls in flight to D (sdft_launch.hpp).
 *      On a streaming plan every thread ends with griddepcontrol.wait: a call completes only after its
 *      predecessor has, so work queued behind the calls in the ordinary way still sees all of them finished
 *      (a plan that never streams gets the same from the wait at the top of each call).
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
  PhaseSource<F> phase;    // where chunk-start phases come from (sdft_schedule.cuh); also carries the mirror-cell map
  const cx<F>* acc_in;     // (channels, cells)
  cx<F>* acc_out;
  cx<F>* totals;           // (channels, nblocks, groups, WC) aggregate of each block item
  cx<F>* prefix;           // (channels, nblocks, groups, WC) inclusive prefix after each block item
  unsigned* flags;         // (channels, nblocks, groups): 2*epoch = aggregate published, 2*epoch+1 = prefix published
  unsigned* ticket;        // work ticket of this call's scratch slot (rearmed by the last ticket of the launch)
  unsigned* error;         // spin-wait timeout flag of the plan
  unsigned* sync;          // this call's hand-over counters: [0] history pieces written, [1] accumulator rows written
  const unsigned* prev_sync;   // the previous call's counters ...
  unsigned prev_hist_target;   // ... and the values they reach once that call has handed over history / accumulators
  unsigned prev_acc_target;
  unsigned flow;           // 1: streaming call, may overlap the previous call (see the header comment); 0: serial
  unsigned handover;       // 1: plan in streaming mode -- keep the hand-over and completion counters (a later call may
                           // poll them); 0: every call on the plan is serial, nobody ever looks
  unsigned* finished;      // CTAs of this call that are through (this call's slot); the last one rearms it and ...
  unsigned* completed;     // ... bumps the plan's count of completed calls (calls complete in order)
  unsigned completed_target;   // streaming: calls that must have completed before this one may touch its slot (mod 2^32)
  unsigned wait_completed;     // 0 for the first `depth` calls on fresh rings: nothing to wait for
  unsigned epoch;
  unsigned total_blocks;   // nblocks * channels * groups: the tickets of this chain set
  unsigned finish_blocks;  // CTAs of the whole launch (this chain set, plus the other one of a mixed launch)
  unsigned nblocks;        // block items per chain: ceil(nchunks / warps per CTA)
  unsigned channels;
  unsigned m;
  unsigned cells;
  cx<F>* out;              // (channels, n, roi_count) or nullptr
  size_t out_channel_stride;
  unsigned roi_first;      // region of interest: rows hold bins [roi_first, roi_first + roi_count) only (sdft.h:137-143)
  unsigned roi_count;      // m for the whole spectrum
  const F* syn_ab;         // (m, 2) synthesis weights of Re / Im of every bin with the window folded in, EMIT_SYNTH only
  F* part;                 // (channels, groups, n) per-group partial sums of the fused synthesis
  unsigned groups;
  unsigned bin_base;       // this launch covers bins [bin_base, bin_end) in `groups` warp-wide groups: everything,
  unsigned bin_end;        // or the wide body / the narrow tail of a split call
  unsigned roll_hist;      // 1: this launch writes the next history, every CTA a slice (exactly one launch of a call does)
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
/* completion order: a thread that has passed this has seen the previous kernel of the stream complete (a no-op
 * for the second time and for kernels launched without programmatic serialization) */
__device__ __forceinline__ void grid_dependency_wait()
{
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
/* last instruction of every warp of the scan kernel: keep the order of completion, then count this warp out; the
 * last warp of the last CTA rearms the slot's counter and bumps the plan's count of completed calls */
__device__ __forceinline__ void warp_finish(unsigned* s_warps_done, unsigned nwarps, unsigned* finished, unsigned total_blocks,
                                            unsigned* completed, unsigned handover)
{
  /* a plan that never streams needs none of this: each of its calls waits for its predecessor at the top, so it
   * cannot complete before it */
  if (!handover) return;
  grid_dependency_wait();
  if (!completed) return;       // the body launch of a split call: the tail launch counts the call
  __syncwarp();
  if ((threadIdx.x & 31u) == 0u)
  {
    if (atomicAdd(s_warps_done, 1u) == nwarps - 1u)
    {
      __threadfence();
      if (atomicAdd(finished, 1u) == total_blocks - 1u)
      {
        *finished = 0;
        __threadfence();
        asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(completed), "r"(1u) : "memory");
      }
    }
  }
}
__device__ __forceinline__ void red_release_add_u32(unsigned* p, unsigned v)
{
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
/* one thread: wait until the monotonic counter has reached `target` (wrap-safe), acquire what its writers released */
__device__ __forceinline__ void wait_counter(const unsigned* ctr, unsigned target, unsigned* error)
{
  unsigned spins = 0;
  unsigned long long t_start = 0;
  while ((int)(ld_acquire_u32(ctr) - target) < 0)
  {
    if (++spins < 64u) continue;
    __nanosleep(100);
    if (t_start == 0) t_start = global_timer_ns();
    else if (global_timer_ns() - t_start > kSpinLimitNs)
    {
      atomicExch(error, 1u);
      break;
    }
  }
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
    const T oldest = (t < period) ? __ldcg(ho + t) : x[t - period];     // history through L2: a streaming call reads
                                                                       // what another kernel wrote a moment ago
    const T diff = newest - oldest;
    sdelta[i] = (F)diff * a.scale;
  }
}
/* the history the next call starts from; entries are dealt out over ALL CTAs of the channel (a one-sample call at
 * m = 4096 still moves 8192 entries: one warp alone needs 40 us for that, the call's 69 CTAs together 2 us) */
template <typename T, typename F>
__device__ __forceinline__ void roll_history(const ChainArgs<F>& a, unsigned ch, unsigned slice, unsigned nslices)
{
  const unsigned period = a.sched.period;
  const T* x = (const T*)a.samples + (size_t)ch * a.sample_stride;
  const T* ho = (const T*)a.hist_old + (size_t)ch * period;
  T* hn = (T*)a.hist_new + (size_t)ch * period;
  for (unsigned i = slice * blockDim.x + threadIdx.x; i < period; i += nslices * blockDim.x)
  {
    const unsigned long long pos = a.sched.n + i;   // position inside history || samples
    hn[i] = (pos < period) ? __ldcg(ho + pos) : x[pos - period];
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

/* the plan's accumulators at the start of this call (one warp, this lane's cells e0 .. e0+CPL-1).  A streaming
 * call may get here while the previous call is still adding up: wait for its accumulator rows first. */
template <typename F, int CPL>
__device__ __forceinline__ void load_plan_acc(const ChainArgs<F>& a, unsigned ch, int e0, unsigned lane, cx<F>* acc, bool wait)
{
  if (wait)
  {
    if (lane == 0) wait_counter(a.prev_sync + 1, a.prev_acc_target, a.error);
    __syncwarp();
  }
  const cx<F>* ai = a.acc_in + (size_t)ch * a.cells;
#pragma unroll
  for (int b = 0; b < CPL; ++b)
  {
    const int e = e0 + b;
    if (e >= 0 && e < (int)a.cells) acc[b] = load_l2<F>(ai + e);
    else { acc[b].r = (F)0; acc[b].i = (F)0; }
  }
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
                                          unsigned ch, int e0, cx<F>* stage, unsigned stage_rows, cx<F>* acc, unsigned trace_slot)
{
  /* `acc` arrives holding the plan's accumulators when the call is serial (loaded before the walk so that the
   * latency hides behind it); a streaming call fetches them only if the walk ends at the chain's start */
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
      if (lane == 0) atomicExch(a.error, 1u);
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
    if (a.flow) load_plan_acc<F, G::CPL>(a, ch, e0, lane, acc, true);
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

/* one time step of the fused synthesis for one lane: advance the replay, weigh the (un-windowed) spectrum.
 * SLIDE (double fast mode): acc holds z = aux + delta_next in every cell, so the constant is taken out
 * again through the sum of the real-part weights (SynthLane::wsum). */
template <bool RESTART, typename F, bool SLIDE, bool FUSED, typename Lane, typename Syn>
__device__ __forceinline__ F synth_step_impl(Lane& L, const Syn& syn, const F* sdelta, unsigned i, const cx<F>* restart)
{
  if constexpr (SLIDE)
  {
    const F d_next = sdelta[i + 1];
#pragma unroll
    for (int b = 0; b < Lane::G::CPL; ++b) L.acc[b] = Arith<F>::horner(L.acc[b], L.tw[b], d_next);
    return fma(-d_next, syn.wsum, syn.weigh(L.acc));
  }
  else
  {
    cx<F> x[Lane::G::CPL];
    L.template advance<RESTART, FUSED>(sdelta[i], restart, x);
    return syn.weigh(x);
  }
}

/* everything a CTA does once it holds its work ticket (phases 0, A, B, C of the header comment), for one
 * instantiation of the lane engine.  `a.total_blocks` is what the TICKETS of this chain set run to; the kernels
 * below take the ticket and decide which chain set it belongs to. */
template <typename F, int WINDOW, bool VEC, int EMIT, int MODE, int GEO>
__device__ __forceinline__ void scan_emit_body(const ChainArgs<F>& a, const unsigned ticket, unsigned char* smem_raw,
                                               unsigned* s_warps_done_ptr)
{
  typedef EmitGeo<F, WINDOW, GEO> G;
  typedef Arith<F> A;
  constexpr bool SLIDE = IsSlide<F, MODE>::value != 0;     // double fast mode
  constexpr bool FUSED = false;                             // stages keep the reference's roundings in every mode
  constexpr bool DPTOTALS = (MODE == MODE_FAST) && !SLIDE;  // float fast mode: chunk totals on the FP64 pipe
  typedef StageOps<F, FUSED> S;
  /* dynamic shared memory, sized by the launch (scan_smem_bytes): deltas of the CTA's chunks, their
   * totals, the carry at the CTA's first chunk */
  const unsigned nwarps = blockDim.x >> 5;
  F* sdelta_all = reinterpret_cast<F*>(smem_raw);
  cx<F>* stot_all = reinterpret_cast<cx<F>*>(smem_raw + (size_t)nwarps * (a.sched.chunk + kDeltaPad) * sizeof(F));
  cx<F>* scarry = stot_all + (size_t)nwarps * G::WC;
  cx<F>* sstage = scarry + G::WC;                            // look-back staging, a.stage_rows rows
#define stot(u) (stot_all + (size_t)(u) * G::WC)
#define s_warps_done (*s_warps_done_ptr)
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
  if (a.roll_hist)
  {
    if (a.td_double) roll_history<double, F>(a, ch, jb * a.groups + group, a.nblocks * a.groups);
    else roll_history<float, F>(a, ch, jb * a.groups + group, a.nblocks * a.groups);
    if (a.handover)
    {
      /* hand-over: this CTA's piece of the next call's history is written */
      __syncthreads();
      if (threadIdx.x == 0) red_release_add_u32(a.sync, 1u);
    }
  }
  __syncwarp();
  SDFT_B200_STAMP(1);   // deltas in shared memory

  EmitLane<F, WINDOW, VEC, GEO> L;
  const bool synth = (EMIT == EMIT_SYNTH_UNIT || EMIT == EMIT_SYNTH);       // the fused synthesis always sees every bin
  const unsigned roi_first = synth ? 0u : a.roi_first, roi_end = synth ? a.m : a.roi_first + a.roi_count;
  const int e0 = L.setup(group, lane, a.m, roi_first, roi_end, a.bin_base, a.bin_end);
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
      /* the starting phase is fetched now (one root of unity per cell) so that its latency hides under the sum */
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
        L.ph[b] = live[b] ? phase_at<F>(a.phase, e0 + b, cs.cursor0, L.tw[b]) : zero;
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
#pragma unroll
      for (int b = 0; b < G::CPL; ++b) tot[b] = X::cmul(L.ph[b], tot[b]);
    }
    else if constexpr (DPTOTALS)
    {
      /* float, fast mode: total = P_start * sum_i tw^i delta_i by Horner in DOUBLE.  The products of two
       * floats are exact in double, so this is the chunk's sum with the float twiddle's own (systematic)
       * error and without the float recurrence's per-step rounding noise -- closer to the exact sum than
       * the reference's own float accumulation, off the strict replay by < 1e-6 of a term (the gate is
       * 1e-4).  It moves the totals from the FP32 pipe, which bounds the float kernel, to the idle FP64
       * pipe.  P_start is the bit-exact table phase; the replay (phase C) stays strict. */
      cx<double> h[G::CPL], w[G::CPL];
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        h[b].r = 0.0; h[b].i = 0.0;
        w[b].r = (double)L.tw[b].r; w[b].i = (double)L.tw[b].i;
        L.ph[b] = live[b] ? phase_at<F>(a.phase, e0 + b, cs.cursor0, L.tw[b]) : zero;
      }
#pragma unroll 2
      for (int i = (int)cs.len - 1; i >= 0; --i)
      {
        const double d = (double)sdelta[i];
#pragma unroll
        for (int b = 0; b < G::CPL; ++b) h[b] = Arith<double>::horner(h[b], w[b], d);
      }
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
      {
        const double pr = (double)L.ph[b].r, pi = (double)L.ph[b].i;
        tot[b].r = (F)(pr * h[b].r - pi * h[b].i);
        tot[b].i = (F)(pr * h[b].i + pi * h[b].r);
      }
    }
    else
    {
#pragma unroll
      for (int b = 0; b < G::CPL; ++b)
        L.ph[b] = live[b] ? phase_at<F>(a.phase, e0 + b, cs.cursor0, L.tw[b]) : zero;
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
    if (!a.flow || jb == 0) load_plan_acc<F, G::CPL>(a, ch, e0, lane, carry, a.flow != 0);
    if (jb > 0) look_back<F, GEO>(a, item, item_stride, jb, lane, ch, e0, sstage, a.stage_rows, carry, ticket);
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
      if constexpr (G::HALO == 0)
      {
        /* halo-free geometry (boxcar rows, fused synthesis on a plan of ANY window): no warp owns mirror
         * cells 0 and 1, and cells m+2, m+3 only when m is not a multiple of the warp width.  A later
         * windowed call reads all four as its carry, so they are written here as what they are by
         * construction: the exact (conjugate) copy of their source bin. */
#pragma unroll
        for (int b = 0; b < G::CPL; ++b)
        {
          const int e = e0 + b;
          if (e < 2 || e >= (int)a.m + 2) continue;
          ao[e] = agg[b];
#pragma unroll
          for (int q = 0; q < 4; ++q)
            if (a.phase.mir_src[q] == e)
            {
              cx<F> v = agg[b];
              if (a.phase.mir_conj[q]) v.i = -v.i;
              ao[a.phase.mir_cell[q]] = v;
            }
        }
      }
      else
      {
        /* every cell is written by the launch that OWNS its bin (halo cells are recomputed, not owned): the two
         * launches of a split call add their carries up in different orders */
        const int own_lo = a.bin_base == 0 ? 0 : (int)a.bin_base + 2;
        const int own_hi = a.bin_end == a.m ? (int)a.cells : (int)a.bin_end + 2;
#pragma unroll
        for (int b = 0; b < G::CPL; ++b)
          if (live[b] && e0 + b >= own_lo && e0 + b < own_hi) ao[e0 + b] = agg[b];
      }
      if (a.handover)
      {
        /* hand-over: this chain's accumulators are in place for the next call */
        __syncwarp();
        if (lane == 0) red_release_add_u32(a.sync + 1, 1u);
      }
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
  if (!valid) { warp_finish(&s_warps_done, nwarps, a.finished, a.finish_blocks, a.completed, a.handover); return; }

  /* ---- phase C: replay from the carry and stream the rows out ---- */
  if (EMIT == EMIT_ROWS)
  {
    /* groups without a bin inside the region of interest have done their share (the carries): no rows */
    if (a.bin_base + group * (unsigned)G::SPAN >= roi_end || a.bin_base + (group + 1u) * (unsigned)G::SPAN <= roi_first)
    {
      warp_finish(&s_warps_done, nwarps, a.finished, a.finish_blocks, a.completed, a.handover);
      return;
    }
    const size_t row_stride = a.roi_count;
    L.dst = a.out + (size_t)ch * a.out_channel_stride + (size_t)cs.t0 * row_stride + ((long long)e0 - 2 - (long long)roi_first);
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
        L.ph[b] = live[b] ? phase_at<F>(a.phase, e0 + b, cs.cursor0, L.tw[b]) : zero;
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
        for (int b = 0; b < G::CPL; ++b) restart[b] = live[b] ? phase_restart<F>(a.phase, e0 + b) : zero;
        L.template step<true, FUSED>(sdelta[body], restart, a.win, row_stride);
      }
    }
  }
  else if (EMIT == EMIT_SYNTH_UNIT || EMIT == EMIT_SYNTH)
  {
    /* fused synthesis: the rows never leave the registers (see SynthLane) */
    typedef SynthLane<F, G::CPL, EMIT == EMIT_SYNTH_UNIT> Y;
    Y syn;
    syn.setup(a.syn_ab, e0, L.ok);
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
        L.ph[b] = live[b] ? phase_at<F>(a.phase, e0 + b, cs.cursor0, L.tw[b]) : zero;
        restart[b] = live[b] ? phase_restart<F>(a.phase, e0 + b) : zero;
      }
    }
    /* eight steps per reduction while they last, then single steps; no branch encloses a shuffle */
    const unsigned body = (SLIDE || !cs.wraps) ? cs.len : cs.len - 1;
    unsigned i = 0;
    for (; i + 8 <= body; i += 8)
    {
#pragma unroll
      for (unsigned u = 0; u < 8; ++u) syn.p[u] = synth_step_impl<false, F, SLIDE, FUSED>(L, syn, sdelta, i + u, restart);
      const F total = syn.reduce8(lane);
      if (writer) pdst[i + slot] = total;
    }
    for (; i < body; ++i)
    {
      const F total = Y::warp_sum(synth_step_impl<false, F, SLIDE, FUSED>(L, syn, sdelta, i, restart));
      if (lane == 0) pdst[i] = total;
    }
    if (!SLIDE && cs.wraps)
    {
      const F total = Y::warp_sum(synth_step_impl<true, F, SLIDE, FUSED>(L, syn, sdelta, body, restart));
      if (lane == 0) pdst[body] = total;
    }
  }
  SDFT_B200_STAMP(6);   // warp 0 finished its rows
  warp_finish(&s_warps_done, nwarps, a.finished, a.finish_blocks, a.completed, a.handover);   // completes only after the call before it
}
#undef stot
#undef s_warps_done

/* register budget of the scan kernels: 2 CTAs of 8 warps per SM (128 registers) by default; -DSDFT_B200_MAXNREG=n
 * builds a variant with an explicit cap instead (tools/build_variants.py; measured: the default is the best) */
#if defined(SDFT_B200_MAXNREG)
#define SDFT_B200_SCAN_BOUNDS __maxnreg__(SDFT_B200_MAXNREG)
#else
#define SDFT_B200_SCAN_BOUNDS __launch_bounds__(kScanWarps * 32, 2)
#endif

/* one call = one chain set = one launch */
template <typename F, int WINDOW, bool VEC, int EMIT, int MODE, int GEO>
__global__ void SDFT_B200_SCAN_BOUNDS scan_emit_kernel(const ChainArgs<F> a)
{
  extern __shared__ __align__(32) unsigned char smem_raw[];
  __shared__ unsigned s_ticket;
  __shared__ unsigned s_done;
  /* programmatic dependent launch (see launch_chain): nothing of the previous kernel in the stream may be
   * read or overwritten before it has completed; dependents of THIS kernel may start filling SMs as
   * soon as every CTA of it has got this far */
  if (!a.flow) grid_dependency_wait();
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0)
  {
    s_done = 0;
    if (a.flow && a.wait_completed) wait_counter(a.completed, a.completed_target, a.error);   // the slot's previous user is through
    const unsigned t = atomicAdd(a.ticket, 1u);
    if (t == a.total_blocks - 1) *a.ticket = 0;   // last ticket of the launch: rearm the slot for its next call
    s_ticket = t;
    if (a.flow)
    {
      /* streaming: the previous call may still be writing the history this CTA reads -- for the deltas of
       * the call's first 2m samples (its first chunk is the earliest) or for rolling a short call's history */
      const unsigned jb0 = t / (a.channels * a.groups);
      const unsigned first = jb0 * (blockDim.x >> 5);
      const bool reads_hist = (first < a.sched.nchunks && chunk_span(a.sched, first).t0 < a.sched.period) ||
                              (a.roll_hist && a.sched.n < a.sched.period);
      if (reads_hist) wait_counter(a.prev_sync, a.prev_hist_target, a.error);
    }
  }
  __syncthreads();
  scan_emit_body<F, WINDOW, VEC, EMIT, MODE, GEO>(a, s_ticket, smem_raw, &s_done);
}

/* one call = TWO chain sets in one launch: a wide body over the full warp groups and a narrow tail over the
 * remaining bins (sdft_launch.hpp explains when).  Bins are independent, so the two sets share nothing but the
 * samples, the state buffers and the hand-over counters; what they must share is the MACHINE -- a tail launched on
 * its own streams 512-byte row segments from a handful of warps per SM and takes as long as a tenth of the body.
 * Every `every`-th ticket goes to the tail, the others to the body, both in their own order: a CTA still waits
 * only for CTAs of its own set with smaller tickets, all of which have started.  Serial calls only. */
template <typename F, int WINDOW, bool VEC, int EMIT, int MODE>
__global__ void SDFT_B200_SCAN_BOUNDS scan_emit_mixed_kernel(const ChainArgs<F> body, const ChainArgs<F> tail, const unsigned every)
{
  extern __shared__ __align__(32) unsigned char smem_raw[];
  __shared__ unsigned s_ticket;
  __shared__ unsigned s_done;
  grid_dependency_wait();
  asm volatile("griddepcontrol.launch_dependents;");
  if (threadIdx.x == 0)
  {
    s_done = 0;
    const unsigned t = atomicAdd(body.ticket, 1u);
    if (t == body.total_blocks + tail.total_blocks - 1) *body.ticket = 0;
    s_ticket = t;
  }
  __syncthreads();
  const MixedTicket mt = mixed_ticket(s_ticket, every, tail.total_blocks);     // sdft_chunks.hpp
  if (mt.is_tail) scan_emit_body<F, WINDOW, VEC, EMIT, MODE, GEO_NARROW>(tail, mt.local, smem_raw, &s_done);
  else scan_emit_body<F, WINDOW, VEC, EMIT, MODE, GEO_WIDE>(body, mt.local, smem_raw, &s_done);
}

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


}  // namespace sdftb200
