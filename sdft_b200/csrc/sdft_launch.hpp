/*
 * sdft_launch.hpp -- per-call launch geometry (chunk length, CTA width, warp geometry) and the device-side passes: analysis_chained, synthesis_device.
 * Host side of libsdft_b200.so; included by sdft_b200.cu only (one translation unit).
 */
#pragma once

#include "sdft_plan.hpp"

namespace
{

/* ------------------------------------------------------------------------------------------------
 * device-side passes
 * ---------------------------------------------------------------------------------------------- */
void prof_mark(Plan* p, int which)
{
  if (!p->profiling) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
  cudaEventRecord(e, p->stream);
  p->prof_events[which].push_back(e);
}

/* bins per warp-wide group of a geometry: the warp's cells minus the halo it recomputes on either side (none for
 * the boxcar window, and none in the fused synthesis, which folds the window into its weights: no taps) */
unsigned span_of(const Plan* p, int geo, bool no_halo)
{
  unsigned wc, halo;
  if (p->fd == kF32)
  {
    wc = geo == GEO_WIDE ? (unsigned)Geo<float, GEO_WIDE>::WC : (unsigned)Geo<float, GEO_NARROW>::WC;
    halo = (unsigned)Geo<float, GEO_WIDE>::GROUP;
  }
  else
  {
    wc = geo == GEO_WIDE ? (unsigned)Geo<double, GEO_WIDE>::WC : (unsigned)Geo<double, GEO_NARROW>::WC;
    halo = (unsigned)Geo<double, GEO_WIDE>::GROUP;
  }
  if (p->window == 0 || no_halo) halo = 0;
  return wc - 2 * halo;
}

/* warp-wide groups of bins the plan's m bins are cut into, for either warp geometry */
unsigned groups_for(const Plan* p, int geo = GEO_WIDE, bool no_halo = false)
{
  const unsigned span = span_of(p, geo, no_halo);
  return (unsigned)((p->m + span - 1) / span);
}

/* narrow warps for short calls (see Geo<F, GEO_NARROW>); only the default arithmetic modes carry
 * narrow kernels */
constexpr double kNarrowBelow = 1.5e6;   // total wide warp-steps of a call below which narrow warps win (profiles/r01_geo_sweep.md)
int choose_geo(const Plan* p, size_t n)
{
  const bool default_mode = (p->mode == MODE_FAST);
  if (!default_mode) return GEO_WIDE;
  if (p->forced_geo >= 0) return p->forced_geo;
  const double u = (double)n * (double)groups_for(p, GEO_WIDE) * (double)p->channels;
  return u < kNarrowBelow ? GEO_NARROW : GEO_WIDE;
}

/* bins per row handed to / taken from the caller */
size_t row_bins(const Plan* p) { return p->roi_count ? p->roi_count : p->m; }

/* 32-byte group stores need rows that start on a 32-byte boundary and a region of interest cut on group
 * boundaries */
template <typename F>
bool can_vectorize(const Plan* p, const void* out, size_t out_stride)
{
  const size_t g = Geo<F, GEO_WIDE>::GROUP;
  return (row_bins(p) % g == 0) && (p->roi_first % g == 0) && (((uintptr_t)out) % 32 == 0) && (out_stride % g == 0);
}

/* the measured optimum for a call (any multiple of 32) */
unsigned choose_chunk_free(const Plan* p, size_t n, int geo)
{
  if (p->forced_chunk)
  {
    size_t c = (p->forced_chunk / kF0Stride) * kF0Stride;
    if (c < (size_t)kF0Stride) c = kF0Stride;
    if (c > (size_t)kMaxChunk) c = kMaxChunk;
    return (unsigned)c;
  }
  /* Measured on B200 (tools/chunk_sweep.py, profiles/r01_chunk_sweep.md): the best chunk length is a
   * function of the call's total warp-steps U = samples x groups x channels.  Short chunks expose the
   * per-chunk latencies (ticket, table loads, look-back), long chunks leave SMs without work. */
  const double u = (double)n * (double)groups_for(p, GEO_WIDE) * (double)p->channels;
  if (geo == GEO_NARROW)
  {
    /* profiles/r01_geo_sweep.md: narrow warps like longer chunks earlier, but never so long that a chain
     * has fewer than 16 chunks.  Streaming calls overlap their neighbours, so the latency of a chunk's serial
     * steps hides behind other calls' work and the per-chunk overheads decide: 128 from the start
     * (profiles/r02_stream_sweep.md: 6.7 us instead of 7.3-7.8 per 4096-sample call at m = 512) */
    unsigned c = (u < 16384.0) ? 32u : ((u < 30.0e3) ? 64u : 128u);
    if (p->stream_depth > 1 && c < 128u) c = 128u;
    while (c > 32u && (size_t)c * 16 > n) c >>= 1;
    return c;
  }
  if (u < 16384.0) return 32;
  if (u < 100.0e3) return 64;
  if (u < 4.0e6) return 128;
  if (u < 16.0e6) return 256;
  return kAutoChunk;
}

/* chunk length of a call: the measured optimum, rounded up to a multiple of the float phase table's stride so
 * that every chunk but the first of a call starts on a table row (no rotations) */
unsigned choose_chunk(const Plan* p, size_t n, int geo)
{
  unsigned c = choose_chunk_free(p, n, geo);
  const unsigned s = p->f0_stride;
  c = ((c + s - 1) / s) * s;
  if (c > (unsigned)kMaxChunk) c = kMaxChunk;
  return c;
}

template <typename F, int EMIT, int GEO>
void launch_chain_geo(Plan* p, const ChainArgs<F>& a, bool vec, unsigned warps)
{
  const dim3 grid(a.total_blocks);
  const size_t smem = scan_smem_bytes<F, GEO>(warps, a.sched.chunk) + (size_t)a.stage_rows * Geo<F, GEO>::WC * sizeof(cx<F>);
  constexpr int kDefaultMode = (int)MODE_FAST;
  /* programmatic dependent launch: the CTAs of this call may become resident while the previous kernel
   * of the stream drains.  A serial call waits at the top of the kernel (griddepcontrol.wait) until that
   * kernel has completed and flushed, so nothing else about the ordering changes and only the launch latency
   * is hidden; a streaming call (ChainArgs::flow) starts computing at once. */
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = p->pdl ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(warps * 32);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = p->stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define SDFT_CHAIN_CASE(W, MODE)                                                                       \
  case W:                                                                                              \
    if (vec) cudaLaunchKernelEx(&cfg, scan_emit_kernel<F, W, true, EMIT, MODE, GEO>, a);               \
    else cudaLaunchKernelEx(&cfg, scan_emit_kernel<F, W, false, EMIT, MODE, GEO>, a);                  \
    break;
  const int window = (EMIT == EMIT_SYNTH_UNIT || EMIT == EMIT_SYNTH) ? 0 : p->window;   // no taps in the fused synthesis
  if (GEO == GEO_NARROW || p->mode == kDefaultMode)
  {
    /* the narrow geometry exists for the default mode only (choose_geo) */
    switch (window)
    {
      SDFT_CHAIN_CASE(0, kDefaultMode)
      SDFT_CHAIN_CASE(1, kDefaultMode)
      SDFT_CHAIN_CASE(2, kDefaultMode)
      SDFT_CHAIN_CASE(3, kDefaultMode)
    }
  }
  else if constexpr (GEO == GEO_WIDE)
  {
    constexpr int kOtherMode = (kDefaultMode == (int)MODE_FAST) ? (int)MODE_MODULATED : (int)MODE_FAST;
    switch (window)
    {
      SDFT_CHAIN_CASE(0, kOtherMode)
      SDFT_CHAIN_CASE(1, kOtherMode)
      SDFT_CHAIN_CASE(2, kOtherMode)
      SDFT_CHAIN_CASE(3, kOtherMode)
    }
  }
#undef SDFT_CHAIN_CASE
  p->launches++;
}

template <typename F, int EMIT>
void launch_chain(Plan* p, const ChainArgs<F>& a, bool vec, unsigned warps, int geo)
{
  if (geo == GEO_NARROW) launch_chain_geo<F, EMIT, GEO_NARROW>(p, a, vec, warps);
  else launch_chain_geo<F, EMIT, GEO_WIDE>(p, a, vec, warps);
}

/* one launch for the two chain sets of a split call (scan_emit_mixed_kernel): wide body + narrow tail, default
 * arithmetic mode, a window with taps (the boxcar window has no halo and never leaves a short last group) */
template <typename F, int EMIT>
void launch_mixed(Plan* p, const ChainArgs<F>& body, const ChainArgs<F>& tail, unsigned every, bool vec, unsigned warps)
{
  const size_t smem_b = scan_smem_bytes<F, GEO_WIDE>(warps, body.sched.chunk) + (size_t)body.stage_rows * Geo<F, GEO_WIDE>::WC * sizeof(cx<F>);
  const size_t smem_t = scan_smem_bytes<F, GEO_NARROW>(warps, tail.sched.chunk) + (size_t)tail.stage_rows * Geo<F, GEO_NARROW>::WC * sizeof(cx<F>);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = p->pdl ? 1 : 0;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(body.total_blocks + tail.total_blocks);
  cfg.blockDim = dim3(warps * 32);
  cfg.dynamicSmemBytes = smem_b > smem_t ? smem_b : smem_t;
  cfg.stream = p->stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
#define SDFT_MIXED_CASE(W)                                                                                       \
  case W:                                                                                                        \
    if (vec) cudaLaunchKernelEx(&cfg, scan_emit_mixed_kernel<F, W, true, EMIT, MODE_FAST>, body, tail, every);   \
    else cudaLaunchKernelEx(&cfg, scan_emit_mixed_kernel<F, W, false, EMIT, MODE_FAST>, body, tail, every);      \
    break;
  switch (p->window)
  {
    SDFT_MIXED_CASE(1)
    SDFT_MIXED_CASE(2)
    SDFT_MIXED_CASE(3)
  }
#undef SDFT_MIXED_CASE
  p->launches++;
}

/* warps (= consecutive chunks) per scan/emit CTA */
unsigned scan_warps_for(const Plan* p, unsigned chunk, unsigned nchunks, int geo)
{
  /* 4 is the measured optimum: wider CTAs shorten the global chain further but pile their stores onto
   * one SM, narrower ones lengthen the chain; long chains of narrow warps gain a few per cent from 8
   * (tools/narrow_sweep.py) */
  unsigned w = p->forced_warps ? p->forced_warps : ((geo == GEO_NARROW && nchunks >= 128u) ? 8u : 4u);
  if (w > (unsigned)kSmemSamples / chunk) w = (unsigned)kSmemSamples / chunk;
  if (w > (unsigned)kScanWarps) w = kScanWarps;
  if (w > nchunks) w = nchunks;
  if (w < 1) w = 1;
  return w;
}

/* one chain set of a call: the bins [bin_base, bin_end) in `groups` warp-wide groups of one geometry */
struct ScanPart
{
  int geo;
  unsigned bin_base, bin_end, groups;
  unsigned slot_id;
  bool roll_hist;      // writes the next history (the first part of a call)
  unsigned warps;
};

/* production path: the single-pass chained scan/emit kernel (deltas and history are its prologue).
 *
 * A call is one chain set, except: long float calls whose last wide warp group would be mostly empty (m = 2048:
 * the 9th group holds 64 of its 248 bins and still issues a full warp's instructions -- the float kernel is bound
 * by instruction issue, so that is 8 % of the time).  Such a call is split by BINS into a wide body over the full
 * groups and a narrow tail over the remaining bins (half the instructions per step), both in ONE launch
 * (scan_emit_mixed_kernel) so that the tail's few warps share the machine with the body.  Bins are independent
 * (every group has its own carry chain): the two sets share nothing but the samples, the state buffers (each
 * writes the accumulators of the bins it owns) and the hand-over counters. */
template <typename T, typename F>
bool analysis_chained(Plan* p, size_t n, const T* x, size_t x_stride, cx<F>* out, size_t out_stride, F* part = nullptr,
                      const F* syn_ab = nullptr, bool syn_unit = false, bool allow_flow = false)
{
  const unsigned m = (unsigned)p->m;
  const unsigned ch = (unsigned)p->channels;
  const unsigned depth = p->stream_depth;
  const int geo = choose_geo(p, n);
  const double work = (double)n * (double)m * (double)ch;
  /* overlap pays where a call's fixed latencies (ticket, first look-back, drain: ~10 us) are a visible share of
   * it: up to ~2^28 bin-updates (0.7 ms).  Longer calls gain nothing and one shape was measured slower
   * (profiles/r02_mid_sweep.md: m = 4096 float, 2^18 samples per call), so they stay serial. */
  const bool may_flow = depth > 1 && allow_flow && !part && work <= 268435456.0;

  ScanPart parts[2];
  int nparts = 1;
  parts[0] = { geo, 0u, m, groups_for(p, geo, part != nullptr), 0u, true, 0u };
  if (geo == GEO_WIDE && p->fd == kF32 && p->mode == MODE_FAST && !part && !may_flow && p->window != 0 && p->forced_geo < 0 &&
      work >= 67108864.0 && !p->no_split)
  {
    const unsigned span_w = span_of(p, GEO_WIDE, false), span_n = span_of(p, GEO_NARROW, false);
    const unsigned full = m / span_w, rest = m - full * span_w;
    if (full >= 1 && full <= 11 && rest > 0 && rest <= span_n)
    {
      parts[0] = { GEO_WIDE, 0u, full * span_w, full, 0u, true, 0u };
      parts[1] = { GEO_NARROW, full * span_w, m, 1u, 0u, false, 0u };
      nparts = 2;
    }
  }

  const unsigned long long seq = p->calls_issued++;
  const size_t ring = p->history.size();
  const unsigned prev_slot = p->prev_slot;
  unsigned flow = 0, first_slot = 0, hist_signals = 0, acc_signals = 0;
  ChainArgs<F> args[2];
  for (int q = 0; q < nparts; ++q)
  {
    ScanPart& sp = parts[q];
    /* both chain sets of a split call cut time the same way: the tail's CTAs then carry as many steps as the
     * body's and pay the per-CTA latencies (ticket, deltas, look-back) as rarely */
    const unsigned chunk = (q == 0) ? choose_chunk(p, n, sp.geo) : args[0].sched.chunk;
    const Schedule sched = make_schedule(p->cursor, n, m, chunk);
    const size_t wc = (sp.geo == GEO_NARROW) ? (size_t)Geo<F, GEO_NARROW>::WC : (size_t)Geo<F, GEO_WIDE>::WC;
    /* both chain sets of a split call run in one launch: one CTA width */
    unsigned warps = (q == 0) ? scan_warps_for(p, chunk, sched.nchunks, sp.geo) : parts[0].warps;
    if (warps > (unsigned)kSmemSamples / chunk) warps = (unsigned)kSmemSamples / chunk;
    sp.warps = warps;
    const unsigned nblocks = (sched.nchunks + warps - 1) / warps;
    const size_t items = (size_t)ch * nblocks * sp.groups;
    if (items >= (1ull << 31))
    {
      plan_fail(p, SDFT_B200_ERR_ARG, "analysis: call too large for one launch", __FILE__, __LINE__);
      return false;
    }
    /* Streaming (sdft_b200_set_streaming, depth D > 1): a call that produces rows or only updates the state,
     * reads its samples from device memory, is short and needs little scratch may overlap its predecessors
     * (flow = 1: no wait at the top of the kernel, hand-over through counters, see sdft_scan.cuh "Between
     * calls").  Call number s takes scratch slot s % D and first waits until s - D + 1 calls have completed,
     * i.e. until the previous user of its slot is through: at most D calls are in flight, and the state rings
     * (D + 1 entries) never have a writer and a reader of different calls on one entry.  Everything else (fused
     * round trips: their finish kernel sits between the scan kernels anyway; long calls; the default depth 1)
     * is serial: it waits for all earlier work at the top of the kernel and uses the extra slots D (and D + 1
     * for the tail launch of a split call). */
    const size_t scratch_bytes = 2 * items * wc * sizeof(cx<F>);
    if (q == 0) flow = (may_flow && scratch_bytes <= ((size_t)32 << 20)) ? 1u : 0u;
    sp.slot_id = flow ? (unsigned)(seq % depth) : depth + (unsigned)q;
    if (q == 0) first_slot = sp.slot_id;
    Plan::Slot& slot = p->slots[sp.slot_id];
    if (!reserve(p, slot.prefix, items * wc * sizeof(cx<F>))) return false;
    if (!reserve(p, slot.chain_totals, items * wc * sizeof(cx<F>))) return false;
    const size_t flags_before = slot.flags.bytes;
    if (!reserve(p, slot.flags, items * sizeof(unsigned))) return false;
    if (slot.flags.bytes != flags_before || slot.epoch >= 0x7ffffff0u)
    {
      CU_TRY(p, cudaMemsetAsync(slot.flags.ptr, 0, slot.flags.bytes, p->stream));   // a stream operation: serializes
      slot.epoch = 0;
    }
    slot.epoch++;

    ChainArgs<F>& a = args[q];
    a.sched = sched;
    a.samples = x;
    a.sample_stride = x_stride;
    a.hist_old = p->history[p->state_sel];
    a.hist_new = p->history[(p->state_sel + 1) % ring];
    a.td_double = (type_id<T>::value == kF64) ? 1 : 0;
    a.scale = (F)p->prescale;
    a.tw_ext = (const cx<F>*)p->tw_ext;
    a.phase = phase_source<F>(p);
    a.acc_in = (const cx<F>*)p->acc_state[p->state_sel];
    a.acc_out = (cx<F>*)p->acc_state[(p->state_sel + 1) % ring];
    a.prefix = (cx<F>*)slot.prefix.ptr;
    a.totals = (cx<F>*)slot.chain_totals.ptr;
    a.flags = (unsigned*)slot.flags.ptr;
    a.error = p->control;
    a.handover = depth > 1 ? 1u : 0u;
    a.completed = depth > 1 ? p->control + 1 : nullptr;
    a.completed_target = (unsigned)(seq - depth + 1);      // modulo 2^32, compared wrap-safe; no wait for the first calls:
    a.wait_completed = (seq >= depth) ? 1u : 0u;
    a.ticket = p->control + 2 + 4 * sp.slot_id;
    a.sync = p->control + 3 + 4 * first_slot;          // both launches of a split call hand over through one pair of counters
    a.finished = p->control + 5 + 4 * first_slot;      // ... and count their CTAs out together
    a.prev_sync = p->control + 3 + 4 * prev_slot;
    a.prev_hist_target = p->slots[prev_slot].hist_total;
    a.prev_acc_target = p->slots[prev_slot].acc_total;
    a.flow = flow;
    a.epoch = slot.epoch;
    a.total_blocks = (unsigned)items;
    a.finish_blocks = (unsigned)items;
    a.nblocks = nblocks;
    a.channels = ch;
    a.m = m;
    a.cells = (unsigned)p->cells;
    a.out = out;
    a.out_channel_stride = out_stride;
    a.roi_first = (unsigned)p->roi_first;
    a.roi_count = (unsigned)row_bins(p);
    a.syn_ab = syn_ab;
    a.part = part;
    a.groups = sp.groups;
    a.bin_base = sp.bin_base;
    a.bin_end = sp.bin_end;
    a.roll_hist = sp.roll_hist ? 1u : 0u;
    a.stage_rows = (sp.geo == GEO_NARROW) ? scan_stage_rows<F, GEO_NARROW>(warps, chunk) : scan_stage_rows<F, GEO_WIDE>(warps, chunk);
    a.win = make_window_const<F>(p->m, p->window);   // sdft.h:422, :371
    a.trace = nullptr;
#if defined(SDFT_B200_TRACE)
    if (q == 0 && reserve(p, p->trace, items * 8 * sizeof(unsigned long long)))
    {
      a.trace = (unsigned long long*)p->trace.ptr;
      p->trace_items = items;
    }
#endif
    if (sp.roll_hist) hist_signals += nblocks * sp.groups * ch;      // every CTA of the launch hands over its slice
    acc_signals += sp.groups * ch;
  }
  if (part || out) prof_mark(p, 0);
  if (nparts == 2)
  {
    /* the tail set must have the same CTA width as the body (it may have been clipped by its chunk length) */
    if (parts[1].warps != parts[0].warps)
    {
      plan_fail(p, SDFT_B200_ERR_ARG, "analysis: split call with two CTA widths", __FILE__, __LINE__);
      return false;
    }
    const unsigned total = args[0].total_blocks + args[1].total_blocks;
    args[0].finish_blocks = args[1].finish_blocks = total;
    const unsigned every = mixed_every(args[0].total_blocks, args[1].total_blocks);   // every `every`-th ticket is a tail ticket
    p->split_calls++;
    if (out) launch_mixed<F, EMIT_ROWS>(p, args[0], args[1], every, can_vectorize<F>(p, out, out_stride), parts[0].warps);
    else launch_mixed<F, EMIT_NONE>(p, args[0], args[1], every, false, parts[0].warps);
  }
  else if (part)
  {
    if (syn_unit) launch_chain<F, EMIT_SYNTH_UNIT>(p, args[0], false, parts[0].warps, parts[0].geo);   // all imaginary-part weights are zero
    else launch_chain<F, EMIT_SYNTH>(p, args[0], false, parts[0].warps, parts[0].geo);
  }
  else if (out)
  {
    launch_chain<F, EMIT_ROWS>(p, args[0], can_vectorize<F>(p, out, out_stride), parts[0].warps, parts[0].geo);
  }
  else
  {
    launch_chain<F, EMIT_NONE>(p, args[0], false, parts[0].warps, parts[0].geo);
  }
  CU_TRY(p, cudaGetLastError());
  if (part || out) prof_mark(p, 0);
  /* what this call will have handed over once its group-0 CTAs / last block items are through */
  p->slots[first_slot].hist_total += hist_signals;
  p->slots[first_slot].acc_total += acc_signals;
  p->prev_slot = first_slot;
  p->state_sel = (p->state_sel + 1) % ring;
  p->cursor = (size_t)((p->cursor + n) % (2 * (size_t)m));
  return true;
}

/* analysis over n samples per channel, everything on the device.
 * x: (channels, x_stride) samples; out: (channels, out_stride) complex rows or nullptr (state only). */
template <typename T, typename F>
bool analysis_device(Plan* p, size_t n, const T* x, size_t x_stride, cx<F>* out, size_t out_stride, bool allow_flow = false)
{
  if (n == 0) return true;
  return analysis_chained<T, F>(p, n, x, x_stride, out, out_stride, (F*)nullptr, (const F*)nullptr, false, allow_flow);
}

template <typename T, typename F>
bool synthesis_device(Plan* p, size_t n, const cx<F>* dfts, size_t dft_stride, T* y, size_t y_stride)
{
  if (n == 0) return true;
  const unsigned ch = (unsigned)p->channels;
  size_t blocks = (n + kSynthWarps - 1) / kSynthWarps;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const dim3 grid((unsigned)blocks, ch);
  prof_mark(p, 1);
  if (p->latency == 1)
    synth_kernel<T, F, true><<<grid, kSynthWarps * 32, 0, p->stream>>>(dfts, dft_stride, (const cx<F>*)p->tws, y,
                                                                       y_stride, n, (unsigned)row_bins(p), (unsigned)p->roi_first);
  else
    synth_kernel<T, F, false><<<grid, kSynthWarps * 32, 0, p->stream>>>(dfts, dft_stride, (const cx<F>*)p->tws, y,
                                                                        y_stride, n, (unsigned)row_bins(p), (unsigned)p->roi_first);
  prof_mark(p, 1);
  p->launches++;
  CU_TRY(p, cudaGetLastError());
  return true;
}

}  // namespace
