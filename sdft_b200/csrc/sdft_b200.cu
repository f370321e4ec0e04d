/*
 * sdft_b200.cu -- the one translation unit of libsdft_b200.so: the C-ABI declared in include/sdft_b200.h over
 * sdft_plan.hpp (device-resident plan, tables), sdft_launch.hpp (per-call geometry, kernel launches),
 * sdft_calls.hpp (host/device pointer plumbing) and the kernels (sdft_kernels.cuh and the headers it includes).
 *
 * What lives where (reference: struct sdft_plan, c/src/sdft/sdft.h:137-182):
 *   tables   tw_ext[m+4], tws[m], F0[ceil(2m/32)][m+4]         device, written once per plan
 *   state    rings of history[2m] and acc_state[m+4] per channel (two entries, depth + 1 while streaming); the cursor lives on the host and the
 *            modulation phase is a pure function of it (table row + <32 rotations), so it is not stored
 *   scratch  samples, deltas, chunk totals/prefixes/flags of the chained scan, row tiles   device, grow-only
 *
 * No CPU fallback: every entry point either runs the CUDA path or records an error on the plan.
 */
#include "../../include/sdft_b200.h"
#include "sdft_kernels.cuh"
#include "sdft_host_copy.hpp"
#include "sdft_tables.hpp"
#include "sdft_weights.hpp"

#include <vector>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <atomic>

using namespace sdftb200;

#include "sdft_plan.hpp"
#include "sdft_launch.hpp"
#include "sdft_calls.hpp"

/* ------------------------------------------------------------------------------------------------
 * C-ABI
 * ---------------------------------------------------------------------------------------------- */
#define SDFT_B200_DEFINE(SFX, TD, FD, FDX)                                                                      \
  extern "C" sdft_b200_plan_t* sdft_b200_##SFX##_alloc_batch(size_t m, int window, double latency, size_t ch)  \
  {                                                                                                             \
    return plan_create<TD, FD>(m, window, latency, ch);                         \
  }                                                                                                             \
  extern "C" sdft_b200_plan_t* sdft_b200_##SFX##_alloc_custom(size_t m, int window, double latency)            \
  {                                                                                                             \
    return sdft_b200_##SFX##_alloc_batch(m, window, latency, 1);                                                \
  }                                                                                                             \
  extern "C" sdft_b200_plan_t* sdft_b200_##SFX##_alloc(size_t m)                                                \
  {                                                                                                             \
    return sdft_b200_##SFX##_alloc_batch(m, sdft_b200_window_hann, 1.0, 1);                                     \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_free(sdft_b200_plan_t* p) { plan_destroy(p); }                              \
  extern "C" void sdft_b200_##SFX##_reset(sdft_b200_plan_t* p)                                                  \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_reset: plan type mismatch")) { DeviceGuard on_device(p->device); plan_reset<TD, FD>(p); } \
  }                                                                                                             \
  extern "C" size_t sdft_b200_##SFX##_size(const sdft_b200_plan_t* p) { return p ? p->m : 0; }                  \
  extern "C" int sdft_b200_##SFX##_window(const sdft_b200_plan_t* p) { return p ? p->window : 0; }              \
  extern "C" double sdft_b200_##SFX##_latency(const sdft_b200_plan_t* p) { return p ? p->latency : 0; }         \
  extern "C" void sdft_b200_##SFX##_sdft_n(sdft_b200_plan_t* p, size_t n, const TD* x, FDX* d)                  \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_sdft_n: plan type mismatch")) do_sdft<TD, FD>(p, n, x, (cx<FD>*)d);              \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_sdft_hops(sdft_b200_plan_t* p, size_t nhops, size_t hop, const TD* x,       \
                                              FDX* d, size_t hop_stride)                                        \
  {                                                                                                             \
    if (!typed<TD, FD>(p, "sdft_sdft_hops: plan type mismatch")) return;                                        \
    for (size_t h = 0; h < nhops; ++h)                                                                          \
      if (!do_sdft<TD, FD>(p, hop, x + h * hop, (cx<FD>*)d + h * hop_stride)) return;                           \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_sdft_batch(sdft_b200_plan_t* p, size_t n, const TD* x, FDX* d)              \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_sdft_batch: plan type mismatch")) do_sdft<TD, FD>(p, n, x, (cx<FD>*)d);          \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_sdft(sdft_b200_plan_t* p, TD sample, FDX* d)                                \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_sdft: plan type mismatch")) do_sdft<TD, FD>(p, 1, &sample, (cx<FD>*)d);          \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_sdft_nd(sdft_b200_plan_t* p, size_t n, const TD* x, FDX** d)                \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_sdft_nd: plan type mismatch")) do_sdft_nd<TD, FD>(p, n, x, (cx<FD>**)d);         \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_advance(sdft_b200_plan_t* p, size_t n, const TD* x)                         \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_advance: plan type mismatch")) do_advance<TD, FD>(p, n, x);                      \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_isdft_n(sdft_b200_plan_t* p, size_t n, const FDX* d, TD* y)                 \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_isdft_n: plan type mismatch")) do_isdft<TD, FD>(p, n, (const cx<FD>*)d, y);      \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_isdft_batch(sdft_b200_plan_t* p, size_t n, const FDX* d, TD* y)             \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_isdft_batch: plan type mismatch")) do_isdft<TD, FD>(p, n, (const cx<FD>*)d, y);  \
  }                                                                                                             \
  extern "C" TD sdft_b200_##SFX##_isdft(sdft_b200_plan_t* p, const FDX* d)                                      \
  {                                                                                                             \
    TD y = 0;                                                                                                   \
    if (typed<TD, FD>(p, "sdft_isdft: plan type mismatch")) do_isdft<TD, FD>(p, 1, (const cx<FD>*)d, &y);       \
    return y;                                                                                                   \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_isdft_nd(sdft_b200_plan_t* p, size_t n, const FDX** d, TD* y)               \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_isdft_nd: plan type mismatch")) do_isdft_nd<TD, FD>(p, n, (const cx<FD>**)d, y); \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_roundtrip_n(sdft_b200_plan_t* p, size_t n, const TD* in, TD* out)           \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_roundtrip_n: plan type mismatch")) do_roundtrip<TD, FD>(p, n, in, out);          \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_convolve_n(sdft_b200_plan_t* p, size_t n, const FDX* in, FDX* out)          \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_convolve_n: plan type mismatch")) do_convolve<FD>(p, n, (const cx<FD>*)in, (cx<FD>*)out); \
  }                                                                                                             \
  extern "C" void sdft_b200_##SFX##_roundtrip_gain_n(sdft_b200_plan_t* p, size_t n, const TD* in, TD* out,      \
                                                     const FDX* gains)                                          \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_roundtrip_gain_n: plan type mismatch"))                                          \
      do_roundtrip<TD, FD>(p, n, in, out, (const cx<FD>*)gains);                                                \
  }

SDFT_B200_DEFINE(f32f32, float, float, sdft_b200_cf32_t)
SDFT_B200_DEFINE(f32f64, float, double, sdft_b200_cf64_t)
SDFT_B200_DEFINE(f64f32, double, float, sdft_b200_cf32_t)
SDFT_B200_DEFINE(f64f64, double, double, sdft_b200_cf64_t)

extern "C" int sdft_b200_last_error(const sdft_b200_plan_t* p) { return p ? p->status : g_alloc_error; }

extern "C" const char* sdft_b200_last_error_string(const sdft_b200_plan_t* p)
{
  return p ? p->errmsg : g_alloc_errmsg;
}

extern "C" int sdft_b200_synchronize(sdft_b200_plan_t* p)
{
  if (!p) return SDFT_B200_ERR_ARG;
  DeviceGuard on_device(p->device);
  cudaError_t e = cudaStreamSynchronize(p->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->copy_stream);
  unsigned timed_out = 0;
  if (e == cudaSuccess) e = cudaMemcpy(&timed_out, p->control, sizeof(timed_out), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) plan_fail(p, (int)e, "sdft_b200_synchronize", __FILE__, __LINE__);
  else if (timed_out) plan_fail(p, SDFT_B200_ERR_CHAIN, "chained scan: a carry wait timed out", __FILE__, __LINE__);
  return p->status;
}

extern "C" int sdft_b200_set_stream(sdft_b200_plan_t* p, void* cuda_stream)
{
  if (!p) return SDFT_B200_ERR_ARG;
  cudaStream_t next = (cuda_stream == (void*)-1) ? p->own_stream : (cudaStream_t)cuda_stream;
  if (next == p->stream) return p->status;
  DeviceGuard on_device(p->device);
  cudaStreamSynchronize(p->stream);   // work queued on the old stream must not race with the new one
  p->stream = next;
  return p->status;
}

extern "C" int sdft_b200_set_streaming(sdft_b200_plan_t* p, unsigned depth)
{
  if (!p) return SDFT_B200_ERR_ARG;
  if (depth < 1) depth = 1;
  if (depth > 64) depth = 64;
  if (depth == p->stream_depth) return p->status;
  DeviceGuard on_device(p->device);
  if (p->td == kF32 && p->fd == kF32) plan_rings<float, float>(p, depth);
  else if (p->td == kF32) plan_rings<float, double>(p, depth);
  else if (p->fd == kF32) plan_rings<double, float>(p, depth);
  else plan_rings<double, double>(p, depth);
  return p->status;
}

extern "C" int sdft_b200_set_chunk(sdft_b200_plan_t* p, size_t chunk)
{
  if (!p) return SDFT_B200_ERR_ARG;
  p->forced_chunk = chunk;
  return 0;
}

extern "C" int sdft_b200_set_roi(sdft_b200_plan_t* p, size_t first, size_t count)
{
  if (!p || first + count > p->m || (count == 0 && first != 0)) return SDFT_B200_ERR_ARG;
  p->roi_first = first;
  p->roi_count = (count == p->m) ? 0 : count;
  return 0;
}

extern "C" int sdft_b200_set_profiling(sdft_b200_plan_t* p, int on)
{
  if (!p) return SDFT_B200_ERR_ARG;
  p->profiling = on != 0;
  return 0;
}

extern "C" double sdft_b200_kernel_ms(sdft_b200_plan_t* p, int which, unsigned long long* launches)
{
  if (launches) *launches = 0;
  if (!p || which < 0 || which > 1) return 0.0;
  DeviceGuard on_device(p->device);
  std::vector<cudaEvent_t>& ev = p->prof_events[which];
  double total = 0.0;
  if (!ev.empty()) cudaEventSynchronize(ev.back());
  for (size_t i = 0; i + 1 < ev.size(); i += 2)
  {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) total += ms;
    if (launches) ++*launches;
  }
  for (cudaEvent_t e : ev) cudaEventDestroy(e);
  ev.clear();
  cudaGetLastError();
  return total;
}

extern "C" size_t sdft_b200_debug_trace(sdft_b200_plan_t* p, unsigned long long* stamps, size_t max_items)
{
  if (!p || !p->trace.ptr || !stamps) return 0;
  DeviceGuard on_device(p->device);
  cudaStreamSynchronize(p->stream);
  const size_t items = p->trace_items < max_items ? p->trace_items : max_items;
  if (cudaMemcpy(stamps, p->trace.ptr, items * 8 * sizeof(unsigned long long), cudaMemcpyDeviceToHost) != cudaSuccess)
  {
    cudaGetLastError();
    return 0;
  }
  return items;
}

extern "C" size_t sdft_b200_channels(const sdft_b200_plan_t* p) { return p ? p->channels : 0; }
extern "C" size_t sdft_b200_table_bytes(const sdft_b200_plan_t* p) { return p ? p->table_bytes : 0; }
extern "C" int sdft_b200_device(const sdft_b200_plan_t* p) { return p ? p->device : -1; }
extern "C" unsigned long long sdft_b200_launch_count(const sdft_b200_plan_t* p) { return p ? p->launches : 0; }
extern "C" unsigned long long sdft_b200_split_count(const sdft_b200_plan_t* p) { return p ? p->split_calls : 0; }

extern "C" int sdft_b200_get_twiddles(sdft_b200_plan_t* p, void* analysis, void* synthesis)
{
  if (!p) return SDFT_B200_ERR_ARG;
  DeviceGuard on_device(p->device);
  const size_t cbytes = (p->fd == kF32) ? sizeof(cx<float>) : sizeof(cx<double>);
  cudaError_t e = cudaStreamSynchronize(p->stream);
  if (e == cudaSuccess) e = cudaMemcpy(analysis, (char*)p->tw_ext + 2 * cbytes, p->m * cbytes, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(synthesis, p->tws, p->m * cbytes, cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) plan_fail(p, (int)e, "sdft_b200_get_twiddles", __FILE__, __LINE__);
  return p->status;
}

extern "C" int sdft_b200_get_state(sdft_b200_plan_t* p, size_t channel, size_t* cursor, void* history,
                                   void* accumulators, void* phase)
{
  if (!p || channel >= p->channels) return SDFT_B200_ERR_ARG;
  DeviceGuard on_device(p->device);
  const size_t cbytes = (p->fd == kF32) ? sizeof(cx<float>) : sizeof(cx<double>);
  const size_t tbytes = (p->td == kF32) ? sizeof(float) : sizeof(double);
  cudaError_t e = cudaStreamSynchronize(p->stream);
  if (cursor) *cursor = p->cursor;
  if (e == cudaSuccess && history)
    e = cudaMemcpy(history, (char*)p->history[p->state_sel] + channel * 2 * p->m * tbytes, 2 * p->m * tbytes,
                   cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && accumulators)
  {
    e = cudaMemcpy(accumulators, (char*)p->acc_state[p->state_sel] + (channel * p->cells + 2) * cbytes, p->m * cbytes,
                   cudaMemcpyDeviceToHost);
    if (e == cudaSuccess && p->prescale != 1.0)
    {
      /* fast mode keeps the accumulators multiplied by the folded window factor */
      double* a = (double*)accumulators;
      for (size_t i = 0; i < 2 * p->m; ++i) a[i] /= p->prescale;
    }
  }
  if (e == cudaSuccess && phase)
  {
    const unsigned threads = 128, blocks = (unsigned)((p->cells + threads - 1) / threads);
    if (p->fd == kF32)
      phase_at_kernel<float><<<blocks, threads, 0, p->stream>>>((const cx<float>*)p->tw_ext, phase_source<float>(p),
                                                                (cx<float>*)p->phase_scratch, (unsigned)p->cursor);
    else
      phase_at_kernel<double><<<blocks, threads, 0, p->stream>>>((const cx<double>*)p->tw_ext, phase_source<double>(p),
                                                                 (cx<double>*)p->phase_scratch, (unsigned)p->cursor);
    p->launches++;
    e = cudaStreamSynchronize(p->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy(phase, (char*)p->phase_scratch + 2 * cbytes, p->m * cbytes, cudaMemcpyDeviceToHost);
  }
  if (e != cudaSuccess) plan_fail(p, (int)e, "sdft_b200_get_state", __FILE__, __LINE__);
  return p->status;
}

/* import: the counterpart of sdft_b200_get_state (HOST buffers; NULL = leave that part as it is) */
extern "C" int sdft_b200_set_state(sdft_b200_plan_t* p, size_t channel, size_t cursor, const void* history,
                                   const void* accumulators)
{
  if (!p || channel >= p->channels || cursor >= 2 * p->m) return SDFT_B200_ERR_ARG;
  DeviceGuard on_device(p->device);
  const size_t cbytes = (p->fd == kF32) ? sizeof(cx<float>) : sizeof(cx<double>);
  const size_t tbytes = (p->td == kF32) ? sizeof(float) : sizeof(double);
  cudaError_t e = cudaStreamSynchronize(p->stream);
  p->cursor = cursor;          // one cursor per plan: channels of a batch plan advance together
  if (e == cudaSuccess && history)
    e = cudaMemcpy((char*)p->history[p->state_sel] + channel * 2 * p->m * tbytes, history, 2 * p->m * tbytes,
                   cudaMemcpyHostToDevice);
  if (e == cudaSuccess && accumulators)
  {
    /* extended cell layout: bins at cells 2..m+1, mirror cells as conjugates of their source bins
     * (make_mirrors); fast double mode keeps the accumulators multiplied by the folded window factor */
    std::vector<double> cells(2 * p->cells, 0.0);
    auto bin = [&](size_t k, int part) -> double
    {
      return (p->fd == kF32) ? (double)((const float*)accumulators)[2 * k + part] : ((const double*)accumulators)[2 * k + part];
    };
    for (size_t k = 0; k < p->m; ++k)
    {
      cells[2 * (k + 2)] = bin(k, 0) * p->prescale;
      cells[2 * (k + 2) + 1] = bin(k, 1) * p->prescale;
    }
    for (int q = 0; q < 4; ++q)
    {
      const int c = p->mirrors.cell[q], src = p->mirrors.src[q];
      if (src < 0) continue;
      cells[2 * c] = bin((size_t)src, 0) * p->prescale;
      cells[2 * c + 1] = bin((size_t)src, 1) * p->prescale * (p->mirrors.conj[q] ? -1.0 : 1.0);
    }
    char* dst = (char*)p->acc_state[p->state_sel] + channel * p->cells * cbytes;
    if (p->fd == kF32)
    {
      std::vector<float> narrow(cells.begin(), cells.end());
      e = cudaMemcpy(dst, narrow.data(), p->cells * cbytes, cudaMemcpyHostToDevice);
    }
    else
    {
      e = cudaMemcpy(dst, cells.data(), p->cells * cbytes, cudaMemcpyHostToDevice);
    }
  }
  if (e != cudaSuccess) plan_fail(p, (int)e, "sdft_b200_set_state", __FILE__, __LINE__);
  return p->status;
}

/* roofline denominators measured on this board, now: see sdft_peak.cuh.  Everything on a private stream with
 * CUDA events; `reps` timed launches after 2 warm-up launches; the best launch counts (a burst figure, like
 * MEASURED_PEAKS.json's copy) and, through *sustained, the mean over all of them. */
extern "C" double sdft_b200_measure_hbm(int kind, void* device_buffer, size_t bytes, int reps, double* sustained)
{
  if (sustained) *sustained = 0.0;
  if (!device_buffer || bytes < (1u << 20) || reps < 1 || kind < 0 || kind > 2) return 0.0;
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double* sink = nullptr;
  double best = 0.0, total_ms = 0.0;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess &&
            cudaEventCreate(&e1) == cudaSuccess && cudaMalloc(&sink, sizeof(double)) == cudaSuccess;
  /* a copy splits the buffer into a source half and a destination half */
  const size_t span = (kind == PEAK_COPY) ? bytes / 2 : bytes;
  const unsigned long long groups = span / 32;
  double* dst = (double*)device_buffer;
  const double* src = (kind == PEAK_COPY) ? (const double*)((char*)device_buffer + groups * 32) : (const double*)device_buffer;
  const unsigned blocks = 148 * 8;
  for (int r = -2; ok && r < reps; ++r)
  {
    cudaEventRecord(e0, st);
    if (kind == PEAK_STORE) peak_stream_kernel<PEAK_STORE><<<blocks, 256, 0, st>>>(dst, src, groups, sink);
    else if (kind == PEAK_READ) peak_stream_kernel<PEAK_READ><<<blocks, 256, 0, st>>>(dst, src, groups, sink);
    else peak_stream_kernel<PEAK_COPY><<<blocks, 256, 0, st>>>(dst, src, groups, sink);
    cudaEventRecord(e1, st);
    ok = cudaEventSynchronize(e1) == cudaSuccess && cudaGetLastError() == cudaSuccess;
    float ms = 0.f;
    if (ok && r >= 0 && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0.f)
    {
      const double moved = (double)groups * 32.0 * (kind == PEAK_COPY ? 2.0 : 1.0);
      const double gbps = moved / (ms * 1e-3) / 1e9;
      if (gbps > best) best = gbps;
      total_ms += ms;
      if (sustained) *sustained = moved * (double)(r + 1) / (total_ms * 1e-3) / 1e9;
    }
  }
  if (sink) cudaFree(sink);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
  cudaGetLastError();
  return ok ? best : 0.0;
}

/* FP64 issue ceiling: DFMA per second of a pure DFMA loop on every SM (sdft_peak.cuh), best of `reps` */
extern "C" double sdft_b200_measure_dfma(int reps)
{
  cudaStream_t st = nullptr;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  double* sink = nullptr;
  double best = 0.0;
  bool ok = cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) == cudaSuccess && cudaEventCreate(&e0) == cudaSuccess &&
            cudaEventCreate(&e1) == cudaSuccess && cudaMalloc(&sink, sizeof(double)) == cudaSuccess;
  const unsigned blocks = 148 * 8, iters = 1u << 14;
  for (int r = -2; ok && r < reps; ++r)
  {
    cudaEventRecord(e0, st);
    peak_dfma_kernel<<<blocks, 256, 0, st>>>(sink, iters, 0.999999, 1e-9);
    cudaEventRecord(e1, st);
    ok = cudaEventSynchronize(e1) == cudaSuccess && cudaGetLastError() == cudaSuccess;
    float ms = 0.f;
    if (ok && r >= 0 && cudaEventElapsedTime(&ms, e0, e1) == cudaSuccess && ms > 0.f)
    {
      const double rate = (double)blocks * 256.0 * (double)iters * 8.0 / (ms * 1e-3);
      if (rate > best) best = rate;
    }
  }
  if (sink) cudaFree(sink);
  if (e0) cudaEventDestroy(e0);
  if (e1) cudaEventDestroy(e1);
  if (st) cudaStreamDestroy(st);
  cudaGetLastError();
  return ok ? best : 0.0;
}

/* shard planners for C / C++ callers: the same integer arithmetic as sdft_b200/shard.py (time_shards,
 * channel_shards); no device involved */
extern "C" int sdft_b200_time_shard(size_t nsamples, size_t world, size_t dftsize, size_t rank, size_t* begin, size_t* end,
                                    size_t* halo_begin)
{
  if (world == 0 || dftsize == 0 || rank >= world) return SDFT_B200_ERR_ARG;
  const size_t period = 2 * dftsize;
  const size_t periods = (nsamples + period - 1) / period;
  const size_t base = periods / world, extra = periods % world;
  size_t b = 0, e = 0;
  for (size_t r = 0; r <= rank; ++r)
  {
    b = e;
    e = b + (base + (r < extra ? 1 : 0)) * period;
    if (e > nsamples) e = nsamples;
  }
  if (begin) *begin = b;
  if (end) *end = e;
  if (halo_begin) *halo_begin = b > period ? b - period : 0;
  return 0;
}

extern "C" int sdft_b200_channel_shard(size_t channels, size_t world, size_t rank, size_t* begin, size_t* end)
{
  if (world == 0 || rank >= world) return SDFT_B200_ERR_ARG;
  const size_t base = channels / world, extra = channels % world;
  const size_t b = rank * base + (rank < extra ? rank : extra);
  if (begin) *begin = b;
  if (end) *end = b + base + (rank < extra ? 1 : 0);
  return 0;
}

extern "C" void* sdft_b200_host_alloc(size_t bytes)
{
  void* ptr = nullptr;
  if (cudaMallocHost(&ptr, bytes) != cudaSuccess)
  {
    cudaGetLastError();
    return nullptr;
  }
  return ptr;
}

extern "C" void sdft_b200_host_free(void* ptr)
{
  if (ptr) cudaFreeHost(ptr);
}

/* device memory for callers that are plain C / C++ without the CUDA toolkit (the reference's drivers): a hop
 * buffer allocated here instead of malloc keeps the rows on the GPU between sdft_sdft_n and sdft_isdft_n */
extern "C" void* sdft_b200_device_alloc(size_t bytes)
{
  void* ptr = nullptr;
  if (cudaMalloc(&ptr, bytes ? bytes : 1) != cudaSuccess)
  {
    cudaGetLastError();
    return nullptr;
  }
  return ptr;
}

extern "C" void sdft_b200_device_free(void* ptr)
{
  if (ptr) cudaFree(ptr);
}

/* synchronous copy between host and device memory in either direction (the runtime works out which is which);
 * waits for the plan's queued work first when a plan is given.  0 on success. */
extern "C" int sdft_b200_copy(sdft_b200_plan_t* p, void* dst, const void* src, size_t bytes)
{
  if (p)
  {
    DeviceGuard on_device(p->device);
    cudaError_t e = cudaStreamSynchronize(p->stream);
    if (e == cudaSuccess) e = cudaMemcpy(dst, src, bytes, cudaMemcpyDefault);
    if (e != cudaSuccess) plan_fail(p, (int)e, "sdft_b200_copy", __FILE__, __LINE__);
    return p->status;
  }
  const cudaError_t e = cudaMemcpy(dst, src, bytes, cudaMemcpyDefault);
  if (e != cudaSuccess) cudaGetLastError();
  return (int)e;
}

extern "C" const char* sdft_b200_version(void)
{
  return "sdft_b200 0.2 (sm_100a; analysis: single-pass chained scan + emit; synthesis: warp reduction; fused round trip)";
}
