/*
 * sdft_calls.hpp -- host/device pointer plumbing of the entry points: tiling, pinned staging for pageable buffers, row-pointer variants, fused round trip, convolve.
 * Host side of libsdft_b200.so; included by sdft_b200.cu only (one translation unit).
 */
#pragma once

#include "sdft_launch.hpp"

namespace
{

/* ------------------------------------------------------------------------------------------------
 * host/device pointer plumbing
 * ---------------------------------------------------------------------------------------------- */
/* do two byte ranges share an address? */
bool overlaps(const void* a, size_t abytes, const void* b, size_t bbytes)
{
  const uintptr_t a0 = (uintptr_t)a, b0 = (uintptr_t)b;
  return a0 < b0 + bbytes && b0 < a0 + abytes;
}

size_t tile_rows(const Plan* p, size_t n, size_t row_bytes)
{
  size_t rows = p->tile_bytes / (row_bytes * p->channels);
  if (rows < 1) rows = 1;
  if (rows > n) rows = n;
  return rows;
}

/* samples -> device (no-op for device pointers).  Layout (channels, n). */
template <typename T>
const T* stage_samples(Plan* p, size_t n, const T* samples, bool* ok)
{
  *ok = true;
  if (classify(samples) == kDevice) return samples;
  const size_t bytes = p->channels * n * sizeof(T);
  if (!reserve(p, p->samples, bytes)) { *ok = false; return nullptr; }
  if (cudaMemcpyAsync(p->samples.ptr, samples, bytes, cudaMemcpyHostToDevice, p->stream) != cudaSuccess ||
      cudaEventRecord(p->samples_in, p->stream) != cudaSuccess)
  {
    plan_fail(p, (int)cudaGetLastError(), "H2D samples", __FILE__, __LINE__);
    *ok = false;
    return nullptr;
  }
  p->samples_in_flight = true;     // a page-locked source is read asynchronously: see release_samples
  return (const T*)p->samples.ptr;
}

/* a call that returns while its kernels are still queued (device destination) must at least have finished
 * reading the caller's HOST samples: the caller may overwrite them as soon as the call returns */
bool release_samples(Plan* p)
{
  if (!p->samples_in_flight) return true;
  p->samples_in_flight = false;
  CU_TRY(p, cudaEventSynchronize(p->samples_in));
  return true;
}

/* SMALL calls on host buffers (the reference's per-sample sdft_sdft / sdft_isdft, short hops): copy operations,
 * a second stream and the host copy threads cost more than the data is worth.  The plan keeps a pinned,
 * device-visible mailbox; the samples are memcpy'd into it, the analysis kernel reads them and writes its rows
 * there in place (posted PCIe writes, a few KiB), one stream synchronisation, one memcpy out.  A single sample at
 * m = 1000: 22 us instead of 31 through the tiled path. */
constexpr size_t kMailboxSamples = 256;            // per channel
constexpr size_t kMailboxRowBytes = (size_t)256 << 10;

size_t mailbox_rows_offset(size_t sample_bytes) { return (sample_bytes + 255) / 256 * 256; }

template <typename T, typename F>
bool small_sdft(Plan* p, size_t n, const T* samples, cx<F>* dfts)
{
  const size_t m = row_bins(p), ch = p->channels;
  const size_t sbytes = ch * n * sizeof(T), rbytes = ch * n * m * sizeof(cx<F>), off = mailbox_rows_offset(sbytes);
  if (!reserve_mailbox(p, off + rbytes)) return false;
  memcpy(p->mailbox, samples, sbytes);
  cx<F>* rows = (cx<F>*)((char*)p->mailbox + off);
  if (!analysis_device<T, F>(p, n, (const T*)p->mailbox, n, rows, n * m)) return false;
  CU_TRY(p, cudaStreamSynchronize(p->stream));
  memcpy(dfts, rows, rbytes);
  return true;
}

template <typename T, typename F>
bool small_isdft(Plan* p, size_t n, const cx<F>* dfts, T* samples)
{
  const size_t m = row_bins(p), ch = p->channels;
  const size_t sbytes = ch * n * sizeof(T), rbytes = ch * n * m * sizeof(cx<F>), off = mailbox_rows_offset(sbytes);
  if (!reserve_mailbox(p, off + rbytes)) return false;
  if (!reserve(p, p->tile[0], rbytes)) return false;
  cx<F>* rows = (cx<F>*)((char*)p->mailbox + off);
  memcpy(rows, dfts, rbytes);
  /* the rows go to the device by ONE DMA from the pinned mailbox (a warp reading them in place would pay a PCIe
   * round trip per load: measured 30 us for a 16 KiB row); the few samples come back in place */
  CU_TRY(p, cudaMemcpyAsync(p->tile[0].ptr, rows, rbytes, cudaMemcpyHostToDevice, p->stream));
  if (!synthesis_device<T, F>(p, n, (const cx<F>*)p->tile[0].ptr, n * m, (T*)p->mailbox, n)) return false;
  CU_TRY(p, cudaStreamSynchronize(p->stream));
  memcpy(samples, p->mailbox, sbytes);
  return true;
}

template <typename T, typename F>
bool do_sdft(Plan* p, size_t n, const T* samples, cx<F>* dfts)
{
  if (n == 0) return true;
  DeviceGuard on_device(p->device);
  if (p->mailbox_on && n <= kMailboxSamples && p->channels * n * row_bins(p) * sizeof(cx<F>) <= kMailboxRowBytes &&
      classify(samples) != kDevice && classify(dfts) != kDevice)
    return small_sdft<T, F>(p, n, samples, dfts);
  bool ok = true;
  const T* x = stage_samples<T>(p, n, samples, &ok);
  if (!ok) return false;
  const size_t m = row_bins(p), ch = p->channels;   // bins per row: the region of interest

  if (classify(dfts) == kDevice)
  {
    /* device samples AND device rows: nothing is copied, so a streaming plan may let this call overlap the one
     * before it (sdft_b200_set_streaming states what the caller promises in return) */
    return analysis_device<T, F>(p, n, x, n, dfts, n * m, x == samples) && release_samples(p);
  }

  /* host destination: compute row tiles on the device and stream them out, overlapping the
   * device-to-host copy of tile i with the kernels of tile i+1 */
  const size_t row_bytes = m * sizeof(cx<F>);
  const size_t rows = tile_rows(p, n, row_bytes);
  const size_t ntiles = (n + rows - 1) / rows;
  for (int b = 0; b < 2; ++b)
    if (!reserve(p, p->tile[b], ch * rows * row_bytes)) return false;

  auto compute = [&](size_t i) -> bool
  {
    const int b = (int)(i & 1);
    const size_t t0 = i * rows;
    const size_t len = (t0 + rows <= n) ? rows : n - t0;
    if (i >= 2) CU_TRY(p, cudaStreamWaitEvent(p->stream, p->tile_free[b], 0));
    if (!analysis_device<T, F>(p, len, x + t0, n, (cx<F>*)p->tile[b].ptr, len * m)) return false;
    CU_TRY(p, cudaEventRecord(p->tile_ready[b], p->stream));
    return true;
  };
  if (!compute(0)) return false;
  if (classify(dfts) == kHostPageable && !p->driver_pageable)
  {
    /* device tile -> pinned staging (DMA) -> caller's pages (host threads); see HostCopier */
    for (int b = 0; b < 2; ++b)
      if (!reserve_stage(p, b, ch * rows * row_bytes)) return false;
    auto dma = [&](size_t i) -> bool
    {
      const int b = (int)(i & 1);
      const size_t t0 = i * rows;
      const size_t len = (t0 + rows <= n) ? rows : n - t0;
      CU_TRY(p, cudaStreamWaitEvent(p->copy_stream, p->tile_ready[b], 0));
      CU_TRY(p, cudaMemcpyAsync(p->stage[b], p->tile[b].ptr, ch * len * row_bytes, cudaMemcpyDeviceToHost, p->copy_stream));
      CU_TRY(p, cudaEventRecord(p->tile_free[b], p->copy_stream));
      CU_TRY(p, cudaEventRecord(p->stage_done[b], p->copy_stream));
      return true;
    };
    if (!dma(0)) return false;
    std::vector<CopySeg> segs;
    for (size_t i = 0; i < ntiles; ++i)
    {
      const int b = (int)(i & 1);
      const size_t t0 = i * rows;
      const size_t len = (t0 + rows <= n) ? rows : n - t0;
      if (i + 1 < ntiles)
      {
        if (!compute(i + 1)) return false;    // its staging buffer was emptied by the host copy of tile i-1
        if (!dma(i + 1)) return false;
      }
      CU_TRY(p, cudaEventSynchronize(p->stage_done[b]));
      segs.clear();
      for (size_t c = 0; c < ch; ++c)
        segs.push_back({ dfts + (c * n + t0) * m, (const cx<F>*)p->stage[b] + c * len * m, len * row_bytes });
      HostCopier::get().run(segs);
    }
    CU_TRY(p, cudaStreamSynchronize(p->stream));
    return true;
  }
  for (size_t i = 0; i < ntiles; ++i)
  {
    if (i + 1 < ntiles && !compute(i + 1)) return false;
    const int b = (int)(i & 1);
    const size_t t0 = i * rows;
    const size_t len = (t0 + rows <= n) ? rows : n - t0;
    CU_TRY(p, cudaStreamWaitEvent(p->copy_stream, p->tile_ready[b], 0));
    for (size_t c = 0; c < ch; ++c)
    {
      CU_TRY(p, cudaMemcpyAsync(dfts + (c * n + t0) * m, (cx<F>*)p->tile[b].ptr + c * len * m, len * row_bytes,
                                cudaMemcpyDeviceToHost, p->copy_stream));
    }
    CU_TRY(p, cudaEventRecord(p->tile_free[b], p->copy_stream));
  }
  CU_TRY(p, cudaStreamSynchronize(p->copy_stream));
  CU_TRY(p, cudaStreamSynchronize(p->stream));
  p->samples_in_flight = false;
  return true;
}

template <typename T, typename F>
bool do_advance(Plan* p, size_t n, const T* samples)
{
  if (n == 0) return true;
  DeviceGuard on_device(p->device);
  bool ok = true;
  const bool host = classify(samples) != kDevice;
  const T* x = stage_samples<T>(p, n, samples, &ok);
  if (!ok) return false;
  if (!analysis_device<T, F>(p, n, x, n, (cx<F>*)nullptr, 0, !host)) return false;
  if (host) CU_TRY(p, cudaStreamSynchronize(p->stream));
  return true;
}

template <typename T, typename F>
bool do_isdft(Plan* p, size_t n, const cx<F>* dfts, T* samples)
{
  if (n == 0) return true;
  DeviceGuard on_device(p->device);
  if (p->mailbox_on && n <= kMailboxSamples && p->channels * n * row_bins(p) * sizeof(cx<F>) <= kMailboxRowBytes &&
      classify(samples) != kDevice && classify(dfts) != kDevice)
    return small_isdft<T, F>(p, n, dfts, samples);
  const size_t m = row_bins(p), ch = p->channels;   // bins per row: the region of interest
  const bool out_dev = classify(samples) == kDevice;
  T* y = samples;
  if (!out_dev)
  {
    if (!reserve(p, p->synth_out, ch * n * sizeof(T))) return false;
    y = (T*)p->synth_out.ptr;
  }

  if (classify(dfts) == kDevice)
  {
    if (!synthesis_device<T, F>(p, n, dfts, n * m, y, n)) return false;
  }
  else
  {
    const size_t row_bytes = m * sizeof(cx<F>);
    const size_t rows = tile_rows(p, n, row_bytes);
    const size_t ntiles = (n + rows - 1) / rows;
    for (int b = 0; b < 2; ++b)
      if (!reserve(p, p->tile[b], ch * rows * row_bytes)) return false;
    auto upload = [&](size_t i) -> bool
    {
      const int b = (int)(i & 1);
      const size_t t0 = i * rows;
      const size_t len = (t0 + rows <= n) ? rows : n - t0;
      if (i >= 2) CU_TRY(p, cudaStreamWaitEvent(p->copy_stream, p->tile_free[b], 0));
      for (size_t c = 0; c < ch; ++c)
      {
        CU_TRY(p, cudaMemcpyAsync((cx<F>*)p->tile[b].ptr + c * len * m, dfts + (c * n + t0) * m, len * row_bytes,
                                  cudaMemcpyHostToDevice, p->copy_stream));
      }
      CU_TRY(p, cudaEventRecord(p->tile_ready[b], p->copy_stream));
      return true;
    };
    /* make sure earlier work on the compute stream that used the tiles is done */
    CU_TRY(p, cudaStreamSynchronize(p->stream));
    if (classify(dfts) == kHostPageable && !p->driver_pageable)
    {
      /* caller's pages -> pinned staging (host threads) -> device tile (DMA); see HostCopier */
      for (int b = 0; b < 2; ++b)
        if (!reserve_stage(p, b, ch * rows * row_bytes)) return false;
      std::vector<CopySeg> segs;
      auto fill = [&](size_t i)
      {
        const int b = (int)(i & 1);
        const size_t t0 = i * rows;
        const size_t len = (t0 + rows <= n) ? rows : n - t0;
        segs.clear();
        for (size_t c = 0; c < ch; ++c)
          segs.push_back({ (cx<F>*)p->stage[b] + c * len * m, dfts + (c * n + t0) * m, len * row_bytes });
        HostCopier::get().run(segs);
      };
      fill(0);
      for (size_t i = 0; i < ntiles; ++i)
      {
        const int b = (int)(i & 1);
        const size_t t0 = i * rows;
        const size_t len = (t0 + rows <= n) ? rows : n - t0;
        if (i >= 2) CU_TRY(p, cudaStreamWaitEvent(p->copy_stream, p->tile_free[b], 0));
        CU_TRY(p, cudaMemcpyAsync(p->tile[b].ptr, p->stage[b], ch * len * row_bytes, cudaMemcpyHostToDevice, p->copy_stream));
        CU_TRY(p, cudaEventRecord(p->tile_ready[b], p->copy_stream));
        CU_TRY(p, cudaEventRecord(p->stage_done[b], p->copy_stream));
        if (i + 1 < ntiles)
        {
          if (i >= 1) CU_TRY(p, cudaEventSynchronize(p->stage_done[b ^ 1]));   // its previous upload has left the buffer
          fill(i + 1);
        }
        CU_TRY(p, cudaStreamWaitEvent(p->stream, p->tile_ready[b], 0));
        if (!synthesis_device<T, F>(p, len, (const cx<F>*)p->tile[b].ptr, len * m, y + t0, n)) return false;
        CU_TRY(p, cudaEventRecord(p->tile_free[b], p->stream));
      }
    }
    else
    {
    if (!upload(0)) return false;
    for (size_t i = 0; i < ntiles; ++i)
    {
      if (i + 1 < ntiles && !upload(i + 1)) return false;
      const int b = (int)(i & 1);
      const size_t t0 = i * rows;
      const size_t len = (t0 + rows <= n) ? rows : n - t0;
      CU_TRY(p, cudaStreamWaitEvent(p->stream, p->tile_ready[b], 0));
      if (!synthesis_device<T, F>(p, len, (const cx<F>*)p->tile[b].ptr, len * m, y + t0, n)) return false;
      CU_TRY(p, cudaEventRecord(p->tile_free[b], p->stream));
    }
    }
  }
  if (!out_dev)
  {
    CU_TRY(p, cudaMemcpyAsync(samples, y, ch * n * sizeof(T), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(p, cudaStreamSynchronize(p->stream));
  }
  else if (classify(dfts) != kDevice)
  {
    /* device destination: the call returns with its kernels still queued, but the caller's HOST rows have
     * been read completely (same rule as release_samples for the analysis) */
    CU_TRY(p, cudaStreamSynchronize(p->copy_stream));
  }
  return true;
}

/* ------------------------------------------------------------------------------------------------
 * row-pointer variants (sdft.h:622-628, 681-687): `n` row pointers, each a host or a device pointer.
 *
 * No API call per row.  The pointer list is cut into RUNS of rows that follow one another in memory
 * (rows[i+1] == rows[i] + bins -- what callers who slice one matrix into rows hand over, e.g. a ring of hop
 * buffers).  Few long runs: every run is one contiguous call through the same path as sdft_sdft_n /
 * sdft_isdft_n (straight into device rows, tiled + double-buffered DMA for host rows).  Many short runs
 * (really scattered rows): the rows are produced in device tiles and moved by ONE scatter/gather kernel per
 * tile over a device copy of the pointer list (device rows), or by one DMA of the tile to pinned staging and
 * the library's host copy threads (host rows).
 * ---------------------------------------------------------------------------------------------- */
struct RowRun { size_t first, count; };

/* Is this row pointer device memory?  One runtime query per ROW would cost more than the row (0.2 us against the
 * 0.03 us a 16 KiB row takes in HBM), so the device allocation a row was found in is remembered and every further
 * row inside it is answered from that range (cuMemGetAddressRange through the runtime's driver entry point: no
 * link-time dependency on libcuda).  Host rows are still asked one by one: they cross PCIe anyway. */
struct DeviceRange
{
  uintptr_t lo = 0, hi = 0;
  bool contains(const void* ptr)
  {
    const uintptr_t a = (uintptr_t)ptr;
    if (a >= lo && a < hi) return true;
    if (classify(ptr) != kDevice) return false;
    typedef int (*range_fn)(unsigned long long*, size_t*, unsigned long long);
    static range_fn fn = []() -> range_fn
    {
      void* f = nullptr;
      cudaDriverEntryPointQueryResult st;
      if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &f, cudaEnableDefault, &st) != cudaSuccess || st != cudaDriverEntryPointSuccess)
      {
        cudaGetLastError();
        f = nullptr;
      }
      return (range_fn)f;
    }();
    unsigned long long b = 0;
    size_t size = 0;
    if (fn && fn(&b, &size, (unsigned long long)a) == 0 && size > 0)
    {
      lo = (uintptr_t)b;
      hi = lo + size;
    }
    return true;
  }
};

template <typename P>
std::vector<RowRun> row_runs(size_t n, P* const* rows, size_t bins)
{
  std::vector<RowRun> runs;
  size_t first = 0;
  for (size_t i = 1; i <= n; ++i)
    if (i == n || rows[i] != rows[i - 1] + bins)
    {
      runs.push_back({ first, i - first });
      first = i;
    }
  return runs;
}

bool few_long_runs(size_t nruns, size_t n) { return nruns <= 8 || nruns * 256 <= n; }

template <typename T, typename F>
bool do_sdft_nd(Plan* p, size_t n, const T* samples, cx<F>** rows_out)
{
  if (n == 0) return true;
  if (p->channels != 1) { plan_fail(p, SDFT_B200_ERR_ARG, "sdft_nd on a batch plan", __FILE__, __LINE__); return false; }
  DeviceGuard on_device(p->device);
  const size_t m = row_bins(p), row_bytes = m * sizeof(cx<F>);
  const std::vector<RowRun> runs = row_runs<cx<F>>(n, rows_out, m);
  if (few_long_runs(runs.size(), n))
  {
    for (const RowRun& r : runs)
      if (!do_sdft<T, F>(p, r.count, samples + r.first, rows_out[r.first])) return false;
    return true;
  }
  bool ok = true;
  const T* x = stage_samples<T>(p, n, samples, &ok);
  if (!ok) return false;
  const size_t rows = tile_rows(p, n, row_bytes);
  if (!reserve(p, p->tile[0], rows * row_bytes)) return false;
  if (!reserve(p, p->row_ptrs, rows * sizeof(void*))) return false;
  std::vector<CopySeg> segs;
  std::vector<void*> dev_rows;
  DeviceRange device_rows;
  for (size_t t0 = 0; t0 < n; t0 += rows)
  {
    const size_t len = (t0 + rows <= n) ? rows : n - t0;
    if (!analysis_device<T, F>(p, len, x + t0, n, (cx<F>*)p->tile[0].ptr, len * m)) return false;
    /* rows of this tile by memory kind (one query per run start would do; scattered rows are their own runs) */
    dev_rows.assign(len, nullptr);
    size_t ndev = 0;
    for (size_t i = 0; i < len; ++i)
      if (device_rows.contains(rows_out[t0 + i])) { dev_rows[i] = rows_out[t0 + i]; ++ndev; }
    if (ndev)
    {
      CU_TRY(p, cudaMemcpyAsync(p->row_ptrs.ptr, dev_rows.data(), len * sizeof(void*), cudaMemcpyHostToDevice, p->stream));
      size_t blocks = (len * m + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      scatter_rows_kernel<F><<<(unsigned)blocks, 256, 0, p->stream>>>((const cx<F>*)p->tile[0].ptr, (cx<F>* const*)p->row_ptrs.ptr,
                                                                       len, (unsigned)m);
      p->launches++;
      CU_TRY(p, cudaGetLastError());
    }
    if (ndev < len)
    {
      if (!reserve_stage(p, 0, rows * row_bytes)) return false;
      CU_TRY(p, cudaMemcpyAsync(p->stage[0], p->tile[0].ptr, len * row_bytes, cudaMemcpyDeviceToHost, p->stream));
      CU_TRY(p, cudaStreamSynchronize(p->stream));
      segs.clear();
      for (size_t i = 0; i < len; ++i)
        if (!dev_rows[i]) segs.push_back({ rows_out[t0 + i], (const char*)p->stage[0] + i * row_bytes, row_bytes });
      HostCopier::get().run(segs);
    }
    /* the tile and the pointer list are reused in stream order: no host wait for device rows (the pageable pointer
     * list has been staged by the time cudaMemcpyAsync returns) */
  }
  return release_samples(p);
}

template <typename T, typename F>
bool do_isdft_nd(Plan* p, size_t n, const cx<F>** rows_in, T* samples)
{
  if (n == 0) return true;
  if (p->channels != 1) { plan_fail(p, SDFT_B200_ERR_ARG, "isdft_nd on a batch plan", __FILE__, __LINE__); return false; }
  DeviceGuard on_device(p->device);
  const size_t m = row_bins(p), row_bytes = m * sizeof(cx<F>);
  const std::vector<RowRun> runs = row_runs<const cx<F>>(n, rows_in, m);
  if (few_long_runs(runs.size(), n))
  {
    for (const RowRun& r : runs)
      if (!do_isdft<T, F>(p, r.count, rows_in[r.first], samples + r.first)) return false;
    return true;
  }
  const size_t rows = tile_rows(p, n, row_bytes);
  if (!reserve(p, p->tile[0], rows * row_bytes)) return false;
  if (!reserve(p, p->row_ptrs, rows * sizeof(void*))) return false;
  const bool out_dev = classify(samples) == kDevice;
  T* y = samples;
  if (!out_dev)
  {
    if (!reserve(p, p->synth_out, n * sizeof(T))) return false;
    y = (T*)p->synth_out.ptr;
  }
  std::vector<CopySeg> segs;
  std::vector<const void*> dev_rows;
  DeviceRange device_rows;
  for (size_t t0 = 0; t0 < n; t0 += rows)
  {
    const size_t len = (t0 + rows <= n) ? rows : n - t0;
    dev_rows.assign(len, nullptr);
    size_t ndev = 0;
    for (size_t i = 0; i < len; ++i)
      if (device_rows.contains(rows_in[t0 + i])) { dev_rows[i] = rows_in[t0 + i]; ++ndev; }
    if (ndev < len)
    {
      /* host rows: gathered into pinned staging by the copy threads, one DMA into the tile */
      if (!reserve_stage(p, 0, rows * row_bytes)) return false;
      segs.clear();
      for (size_t i = 0; i < len; ++i)
        if (!dev_rows[i]) segs.push_back({ (char*)p->stage[0] + i * row_bytes, rows_in[t0 + i], row_bytes });
      HostCopier::get().run(segs);
      if (ndev == 0)
      {
        CU_TRY(p, cudaMemcpyAsync(p->tile[0].ptr, p->stage[0], len * row_bytes, cudaMemcpyHostToDevice, p->stream));
      }
      else
      {
        for (const CopySeg& sgm : segs)     // mixed tile: only the host rows' slots are uploaded
          CU_TRY(p, cudaMemcpyAsync((char*)p->tile[0].ptr + ((char*)sgm.dst - (char*)p->stage[0]), sgm.dst, row_bytes,
                                    cudaMemcpyHostToDevice, p->stream));
      }
    }
    if (ndev)
    {
      CU_TRY(p, cudaMemcpyAsync(p->row_ptrs.ptr, dev_rows.data(), len * sizeof(void*), cudaMemcpyHostToDevice, p->stream));
      size_t blocks = (len * m + 255) / 256;
      if (blocks > 148 * 16) blocks = 148 * 16;
      gather_rows_kernel<F><<<(unsigned)blocks, 256, 0, p->stream>>>((cx<F>*)p->tile[0].ptr, (const cx<F>* const*)p->row_ptrs.ptr,
                                                                      len, (unsigned)m);
      p->launches++;
      CU_TRY(p, cudaGetLastError());
    }
    if (!synthesis_device<T, F>(p, len, (const cx<F>*)p->tile[0].ptr, len * m, y + t0, n)) return false;
    if (ndev < len) CU_TRY(p, cudaStreamSynchronize(p->stream));     // the staging buffer is refilled by the next tile
  }
  if (!out_dev)
  {
    CU_TRY(p, cudaMemcpyAsync(samples, y, n * sizeof(T), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(p, cudaStreamSynchronize(p->stream));
  }
  return true;
}

/* Synthesis weights with the window folded in: see sdft_weights.hpp */
template <typename F>
bool make_synth_weights(const Plan* p, const std::vector<cx<double>>& v, std::vector<F>& ab)
{
  std::vector<double> vr(p->m), vi(p->m);
  for (size_t k = 0; k < p->m; ++k) { vr[k] = v[k].r; vi[k] = v[k].i; }
  return synth_weights<F>(p->m, p->window, p->mirrors, p->prescale, vr.data(), vi.data(), ab);
}

/* the per-bin factor of sdft_isdft: (-1)^k for latency 1 (exact compare, sdft.h:639), tws[k] otherwise,
 * times an optional spectral gain */
template <typename F>
std::vector<cx<double>> synth_factors(const Plan* p, const cx<F>* gains_host)
{
  const size_t m = p->m;
  std::vector<cx<F>> tw, tws;
  make_tables<F>(m, p->latency, tw, tws);
  std::vector<cx<double>> v(m);
  for (size_t k = 0; k < m; ++k)
  {
    cx<double> t;
    t.r = (double)tws[k].r; t.i = (double)tws[k].i;
    if (p->latency == 1) { t.r = (k & 1) ? -1.0 : 1.0; t.i = 0.0; }
    if (gains_host)
    {
      const double gr = (double)gains_host[k].r, gi = (double)gains_host[k].i;
      const cx<double> u = t;
      t.r = gr * u.r - gi * u.i;
      t.i = gr * u.i + gi * u.r;
    }
    v[k] = t;
  }
  return v;
}

/* analysis -> synthesis in ONE kernel: the rows never exist in memory.  Every warp weighs and reduces
 * its bins per time step (SynthLane), per-group partial sums go to a scratch buffer and a small second
 * kernel adds the groups in order.  Long calls are cut into pieces that bound the scratch.
 * `gains`: spectral processing between analysis and synthesis -- every row is multiplied bin by bin with
 * `gains` before sdft_isdft sees it. */
template <typename T, typename F>
bool do_roundtrip(Plan* p, size_t n, const T* in, T* out, const cx<F>* gains = nullptr)
{
  if (n == 0) return true;
  DeviceGuard on_device(p->device);
  const size_t m = p->m;
  const F* syn_ab = nullptr;
  bool syn_unit = false;
  if (gains || !p->syn_ab_ready)
  {
    std::vector<cx<F>> g;
    if (gains)
    {
      g.resize(m);
      if (classify(gains) == kDevice) CU_TRY(p, cudaMemcpy(g.data(), gains, m * sizeof(cx<F>), cudaMemcpyDeviceToHost));
      else memcpy(g.data(), gains, m * sizeof(cx<F>));
    }
    std::vector<F> ab;
    const bool unit = make_synth_weights<F>(p, synth_factors<F>(p, gains ? g.data() : nullptr), ab);
    Buffer& dst = gains ? p->weights : p->syn_ab;
    if (!reserve(p, dst, 2 * m * sizeof(F))) return false;
    CU_TRY(p, cudaMemcpyAsync(dst.ptr, ab.data(), 2 * m * sizeof(F), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(p, cudaStreamSynchronize(p->stream));   // ab goes out of scope
    if (!gains) { p->syn_ab_ready = true; p->syn_ab_unit = unit; }
    syn_ab = (const F*)dst.ptr;
    syn_unit = unit;
  }
  else
  {
    syn_ab = (const F*)p->syn_ab.ptr;
    syn_unit = p->syn_ab_unit;
  }
  /* in == out (processing a buffer in place) is fine: every piece reads its own samples before its finish
   * kernel writes them; partially overlapping device ranges are not */
  if (in != out && classify(in) == kDevice && classify(out) == kDevice &&
      overlaps(in, p->channels * n * sizeof(T), out, p->channels * n * sizeof(T)))
  {
    plan_fail(p, SDFT_B200_ERR_ARG, "roundtrip: in and out overlap without being the same buffer", __FILE__, __LINE__);
    return false;
  }
  bool ok = true;
  const T* x = stage_samples<T>(p, n, in, &ok);
  if (!ok) return false;
  const size_t ch = p->channels;
  const unsigned max_groups = groups_for(p, GEO_NARROW, true);    // either geometry may be chosen per piece
  const bool out_dev = classify(out) == kDevice;
  T* y = out;
  if (!out_dev)
  {
    if (!reserve(p, p->synth_out, ch * n * sizeof(T))) return false;
    y = (T*)p->synth_out.ptr;
  }
  size_t piece = env_size("SDFT_B200_ROUNDTRIP_PIECE", (size_t)1 << 22);
  {
    /* the partial-sum scratch is (channels, groups, piece): wide batches get shorter pieces so that it stays
     * bounded (1 GiB by default; 512 channels x m = 1024 -> 16 Ki samples per launch, 2e9 bin-updates each) */
    const size_t cap = env_size("SDFT_B200_ROUNDTRIP_SCRATCH_MB", 1024) << 20;
    size_t fit = cap / (ch * max_groups * sizeof(F));
    fit = (fit / 4096) * 4096;
    if (fit < 4096) fit = 4096;
    if (piece > fit) piece = fit;
  }
  if (piece > n) piece = n;
  if (!reserve(p, p->part, ch * max_groups * piece * sizeof(F))) return false;
  for (size_t t0 = 0; t0 < n; t0 += piece)
  {
    const size_t len = (t0 + piece <= n) ? piece : n - t0;
    const unsigned groups = groups_for(p, choose_geo(p, len), true);   // what analysis_chained will use for this piece
    if (!analysis_chained<T, F>(p, len, x + t0, n, (cx<F>*)nullptr, 0, (F*)p->part.ptr, syn_ab, syn_unit)) return false;
    size_t blocks = (len + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    synth_finish_kernel<T, F><<<dim3((unsigned)blocks, (unsigned)ch), 256, 0, p->stream>>>(
        (const F*)p->part.ptr, groups, len, y + t0, n);
    p->launches++;
    CU_TRY(p, cudaGetLastError());
  }
  if (!out_dev)
  {
    CU_TRY(p, cudaMemcpyAsync(out, y, ch * n * sizeof(T), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(p, cudaStreamSynchronize(p->stream));
    p->samples_in_flight = false;
    return true;
  }
  return release_samples(p);
}

/* SDFT.convolve of the reference's Python class (python/src/sdft/sdft.py:146-203): window(rows) / m */
template <typename F>
bool do_convolve(Plan* p, size_t n, const cx<F>* in, cx<F>* out)
{
  if (n == 0) return true;
  DeviceGuard on_device(p->device);
  const size_t m = p->m, ch = p->channels;
  const int need = (p->window == 3) ? 3 : ((p->window == 0) ? 1 : 2);
  if ((int)m < need)
  {
    plan_fail(p, SDFT_B200_ERR_ARG, "convolve: dftsize too small for this window", __FILE__, __LINE__);
    return false;
  }
  const size_t bytes = ch * n * m * sizeof(cx<F>);
  const bool in_dev = classify(in) == kDevice, out_dev = classify(out) == kDevice;
  if (in_dev && out_dev && overlaps(in, bytes, out, bytes))
  {
    /* every output bin reads its neighbours k-2 .. k+2 of the input row: not an in-place operation */
    plan_fail(p, SDFT_B200_ERR_ARG, "convolve: in and out must not overlap in device memory", __FILE__, __LINE__);
    return false;
  }
  const cx<F>* src = in;
  cx<F>* dst = out;
  if (!in_dev)
  {
    if (!reserve(p, p->tile[0], bytes)) return false;
    CU_TRY(p, cudaMemcpyAsync(p->tile[0].ptr, in, bytes, cudaMemcpyHostToDevice, p->stream));
    src = (const cx<F>*)p->tile[0].ptr;
  }
  if (!out_dev)
  {
    if (!reserve(p, p->tile[1], bytes)) return false;
    dst = (cx<F>*)p->tile[1].ptr;
  }
  const F scale = (F)1 / (F)m;
  F c0 = scale, c1 = 0, c2 = 0;
  if (p->window == 1) { c0 = (F)0.5 * scale; c1 = (F)0.25 * scale; }
  if (p->window == 2) { c0 = (F)0.54 * scale; c1 = (F)0.23 * scale; }
  if (p->window == 3) { c0 = (F)0.42 * scale; c1 = (F)0.25 * scale; c2 = (F)0.04 * scale; }
  size_t blocks = (ch * n * m + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  convolve_kernel<F><<<(unsigned)blocks, 256, 0, p->stream>>>(src, dst, ch * n, (unsigned)m, p->window, c0, c1, c2);
  p->launches++;
  CU_TRY(p, cudaGetLastError());
  if (!out_dev)
  {
    CU_TRY(p, cudaMemcpyAsync(out, dst, bytes, cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(p, cudaStreamSynchronize(p->stream));
  }
  return true;
}

template <typename T, typename F>
bool typed(Plan* p, const char* fn)
{
  if (!p) return false;
  if (p->td != type_id<T>::value || p->fd != type_id<F>::value)
  {
    plan_fail(p, SDFT_B200_ERR_TYPE, fn, __FILE__, __LINE__);
    return false;
  }
  return true;
}

}  // namespace
