/*
 * sdft_plan.hpp -- the device-resident plan (struct sdft_b200_plan), error reporting, host-side table generation in the reference's expression order (c/src/sdft/sdft.h:137-182, :413-554).
 * Host side of libsdft_b200.so; included by sdft_b200.cu only (one translation unit).
 */
#pragma once


/* ------------------------------------------------------------------------------------------------
 * plan
 * ---------------------------------------------------------------------------------------------- */
enum TypeId { kF32 = 0, kF64 = 1 };

struct Buffer
{
  void* ptr = nullptr;
  size_t bytes = 0;
};

struct sdft_b200_plan
{
  int td = kF32, fd = kF64;
  size_t m = 0;
  size_t cells = 0;
  int window = 1;
  double latency = 1;
  size_t channels = 1;
  int device = 0;

  cudaStream_t own_stream = nullptr;
  cudaStream_t copy_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaEvent_t samples_in = nullptr;       // recorded after the host-to-device copy of a call's samples
  bool samples_in_flight = false;
  cudaEvent_t tile_ready[2] = { nullptr, nullptr };
  cudaEvent_t tile_free[2] = { nullptr, nullptr };

  size_t cursor = 0;
  size_t roi_first = 0, roi_count = 0;   // region of interest of the rows (sdft.h:137-143); roi_count == 0: all m bins
  size_t forced_chunk = 0;
  unsigned forced_warps = 0;     // SDFT_B200_WARPS: warps per scan/emit CTA (0 = choose per plan geometry)
  int forced_geo = -1;           // SDFT_B200_GEO=wide|narrow: warp geometry (default: per call, choose_geo)
  bool driver_pageable = false;  // SDFT_B200_PAGEABLE=driver: leave pageable buffers to cudaMemcpy (for comparison)
  bool pdl = true;               // SDFT_B200_PDL=0: plain stream-ordered launches
  bool mailbox_on = true;        // SDFT_B200_MAILBOX=0: small host-buffer calls take the tiled path like everything else (for comparison)
  bool no_split = false;         // SDFT_B200_NO_SPLIT=1: never split a float call into wide body + narrow tail (for comparison)
  size_t tile_bytes = 0;
  unsigned long long launches = 0;
  unsigned long long split_calls = 0;    // calls that ran as wide body + narrow tail in one launch (sdft_launch.hpp)

  MirrorMap mirrors;
  int mode = 0;                  // MODE_MODULATED / MODE_FAST (double frequency domain only)
  double prescale = 1.0;         // factor folded into the deltas in fast mode (acc_state is scaled by it)
  void* tw_ext = nullptr;
  void* tws = nullptr;
  void* f0 = nullptr;            // float: (f0_rows, cells) phase table at every f0_stride-th cursor; double: the P[0] row only
  size_t f0_rows = 0;
  unsigned f0_stride = 32;
  void* roots = nullptr;         // double: the 2m roots of unity E[j] = exp(-2 pi i j / 2m) (PhaseSource)
  size_t table_bytes = 0;        // device bytes of tw_ext + tws + f0 + roots

  /* State rings: call e reads entry state_sel and writes entry (state_sel + 1) % size.  Two entries (ping-pong:
   * neighbouring groups share halo cells, and a call reads the old history while it writes the new one) for
   * serial calls; stream_depth + 1 entries while streaming, so that no call in flight writes an entry another
   * call in flight still reads (sdft_launch.hpp: streaming) */
  std::vector<void*> history;      // (channels, 2m) time-domain samples each
  std::vector<void*> acc_state;    // (channels, cells) complex each
  size_t state_sel = 0;
  void* phase_scratch = nullptr;   // cells complex values, introspection only

  Buffer samples, synth_out, tile[2], part, weights;
  Buffer row_ptrs;               // device copy of one tile's row pointers (sdft_sdft_nd / sdft_isdft_nd, scattered rows)
  Buffer syn_ab;                 // fused synthesis: per-bin (A, B) weights with the window folded in (make_synth_weights)
  bool syn_ab_ready = false, syn_ab_unit = false;
  Buffer trace;                  // -DSDFT_B200_TRACE builds: per-CTA phase stamps of the last scan launch
  size_t trace_items = 0;
  void* stage[2] = { nullptr, nullptr };   // pinned host staging for PAGEABLE caller buffers (grow-only)
  void* mailbox = nullptr;       // pinned, device-visible: samples and rows of SMALL host-buffer calls travel through it
                                 // without copy operations (the kernels read and write it in place), see do_sdft
  size_t mailbox_bytes = 0;
  size_t stage_bytes[2] = { 0, 0 };
  cudaEvent_t stage_done[2] = { nullptr, nullptr };
  /* Scratch of the chained scan, one slot per call that may be in flight: inclusive prefixes, block totals,
   * epoch-stamped flags; on the device (control): [0] spin-wait timeout flag of the plan, then per slot a work
   * ticket and the two hand-over counters (history pieces / accumulator rows written), whose running totals the
   * host keeps so that the next call knows what to wait for */
  struct Slot
  {
    Buffer prefix, chain_totals, flags;
    unsigned epoch = 0;
    unsigned hist_total = 0, acc_total = 0;
  };
  std::vector<Slot> slots;       // stream_depth slots for streaming calls (call seq % depth), one for serial calls, one for
                                 // the narrow tail launch of a split call (sdft_launch.hpp)
  unsigned* control = nullptr;   // [0] timeout flag, [1] completed calls, then per slot: ticket, history, accumulators, finished
  unsigned stream_depth = 1;     // calls that may be in flight at once (1: serial; sdft_b200_set_streaming)
  unsigned long long calls_issued = 0;   // analysis calls since the rings were (re)allocated; the device counts them out
                                         // again modulo 2^32 (an endless stream gets there: 8 hours of 4096-sample calls)
  unsigned prev_slot = 0;        // slot of the previous call

  /* optional CUDA-event timing of the dominant kernels (bench.py roofline): [0] analysis emit, [1] synthesis */
  bool profiling = false;
  std::vector<cudaEvent_t> prof_events[2];

  int status = 0;
  char errmsg[256] = "";
};

typedef sdft_b200_plan Plan;

/* ------------------------------------------------------------------------------------------------
 * errors
 * ---------------------------------------------------------------------------------------------- */
namespace
{

/* Every entry point works on the plan's device and leaves the calling thread's current device as it found
 * it: in a multi-GPU process (torch, one thread driving several plans) a library call must not silently
 * move the caller's later allocations and launches to another GPU. */
struct DeviceGuard
{
  int prev = -1;
  bool moved = false;
  explicit DeviceGuard(int device)
  {
    if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; }
    if (prev != device) moved = (cudaSetDevice(device) == cudaSuccess);
  }
  ~DeviceGuard()
  {
    if (moved && prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

enum { SDFT_B200_ERR_TYPE = 10001, SDFT_B200_ERR_ARG = 10002, SDFT_B200_ERR_NODEVICE = 10003, SDFT_B200_ERR_CHAIN = 10004 };

thread_local int g_alloc_error = 0;
thread_local char g_alloc_errmsg[256] = "";

void plan_fail(Plan* p, int code, const char* what, const char* file, int line);

#define CU_TRY(plan, expr)                                         \
  do                                                               \
  {                                                                \
    cudaError_t e__ = (expr);                                      \
    if (e__ != cudaSuccess)                                        \
    {                                                              \
      plan_fail((plan), (int)e__, #expr, __FILE__, __LINE__);      \
      return false;                                                \
    }                                                              \
  } while (0)

template <typename X> struct type_id;
template <> struct type_id<float> { static const int value = kF32; };
template <> struct type_id<double> { static const int value = kF64; };

size_t env_size(const char* name, size_t fallback)
{
  const char* v = getenv(name);
  if (!v || !*v) return fallback;
  char* end = nullptr;
  const unsigned long long x = strtoull(v, &end, 10);
  return (end && end != v) ? (size_t)x : fallback;
}

/* ------------------------------------------------------------------------------------------------
 * plan
 * ---------------------------------------------------------------------------------------------- */
void plan_fail(Plan* p, int code, const char* what, const char* file, int line)
{
  const char* name = (code < 10000) ? cudaGetErrorString((cudaError_t)code) : "sdft_b200 error";
  char msg[256];
  snprintf(msg, sizeof(msg), "%s: %s (%d) at %s:%d", what, name, code, file, line);
  if (p)
  {
    if (p->status == 0)
    {
      p->status = code;
      snprintf(p->errmsg, sizeof(p->errmsg), "%s", msg);
      fprintf(stderr, "[sdft_b200] %s\n", msg);
    }
  }
  else
  {
    g_alloc_error = code;
    snprintf(g_alloc_errmsg, sizeof(g_alloc_errmsg), "%s", msg);
    fprintf(stderr, "[sdft_b200] %s\n", msg);
  }
  if (code < 10000) cudaGetLastError();
}

bool reserve(Plan* p, Buffer& b, size_t bytes)
{
  if (bytes <= b.bytes) return true;
  if (b.ptr)
  {
    CU_TRY(p, cudaStreamSynchronize(p->stream));
    CU_TRY(p, cudaFree(b.ptr));
    b.ptr = nullptr;
    b.bytes = 0;
  }
  const size_t want = bytes + bytes / 8;
  CU_TRY(p, cudaMalloc(&b.ptr, want));
  b.bytes = want;
  return true;
}

enum PtrKind { kHostPageable, kHostPinned, kDevice };

PtrKind classify(const void* ptr)
{
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, ptr) != cudaSuccess)
  {
    cudaGetLastError();
    return kHostPageable;
  }
  if (a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged) return kDevice;
  if (a.type == cudaMemoryTypeHost) return kHostPinned;
  return kHostPageable;
}

/* the mailbox of small host-buffer calls: [samples | rows], grow-only */
bool reserve_mailbox(Plan* p, size_t bytes)
{
  if (bytes <= p->mailbox_bytes) return true;
  if (p->mailbox)
  {
    CU_TRY(p, cudaStreamSynchronize(p->stream));
    CU_TRY(p, cudaFreeHost(p->mailbox));
    p->mailbox = nullptr;
    p->mailbox_bytes = 0;
  }
  CU_TRY(p, cudaMallocHost(&p->mailbox, bytes));
  p->mailbox_bytes = bytes;
  return true;
}

bool reserve_stage(Plan* p, int b, size_t bytes)
{
  if (bytes <= p->stage_bytes[b]) return true;
  if (p->stage[b])
  {
    CU_TRY(p, cudaStreamSynchronize(p->copy_stream));
    CU_TRY(p, cudaFreeHost(p->stage[b]));
    p->stage[b] = nullptr;
    p->stage_bytes[b] = 0;
  }
  CU_TRY(p, cudaMallocHost(&p->stage[b], bytes));
  p->stage_bytes[b] = bytes;
  return true;
}

template <typename F> size_t csize() { return sizeof(cx<F>); }

/* Cursors per row of the float phase table: 32 while the table fits the budget, doubling with m beyond that
 * (up to the longest chunk, 1024), so that the table grows like m^2 / stride only until the budget and stays
 * O(m) per row.  m = 65536: stride 512, 256 rows, 128 MiB.  Past stride 1024 the table itself has to grow; sizes
 * whose table would not fit a quarter of the device are rejected up front (plan_create) with a clear message
 * instead of a cudaMalloc failure deep inside.  SDFT_B200_F0_BUDGET_MB overrides the budget (tests). */
inline unsigned f0_stride_for(size_t m, size_t entry_bytes, size_t budget)
{
  unsigned stride = kF0Stride;
  while (stride < (unsigned)kMaxChunk && ((2 * m + stride - 1) / stride) * (m + 4) * entry_bytes > budget) stride *= 2;
  return stride;
}

template <typename F> PhaseSource<F> phase_source(const sdft_b200_plan* p)
{
  PhaseSource<F> s;
  s.f0 = (const cx<F>*)p->f0;
  s.roots = (const cx<F>*)p->roots;
  s.cells = (unsigned)p->cells;
  s.m = (unsigned)p->m;
  s.period = (unsigned)(2 * p->m);
  s.inv_period = ~0ull / (unsigned long long)(2 * p->m);
  s.stride = p->f0_stride;
  for (int q = 0; q < 4; ++q)
  {
    s.mir_cell[q] = p->mirrors.cell[q];
    s.mir_src[q] = p->mirrors.src[q] < 0 ? -1 : p->mirrors.src[q] + 2;
    s.mir_conj[q] = p->mirrors.conj[q];
  }
  return s;
}

/* -------- plan construction -------- */
/* (re)allocates the state rings and scratch slots for `depth` calls in flight, carrying the current state over */
template <typename T, typename F>
bool plan_rings(Plan* p, unsigned depth)
{
  const size_t m = p->m, cells = p->cells, ch = p->channels;
  const size_t hbytes = ch * 2 * m * sizeof(T), abytes = ch * cells * sizeof(cx<F>);
  if (p->stream) CU_TRY(p, cudaStreamSynchronize(p->stream));
  std::vector<void*> hist(depth + 1, nullptr), acc(depth + 1, nullptr);
  for (unsigned i = 0; i <= depth; ++i)
  {
    cudaError_t e = cudaMalloc(&hist[i], hbytes);
    if (e == cudaSuccess) e = cudaMalloc(&acc[i], abytes);
    if (e == cudaSuccess) e = cudaMemset(hist[i], 0, hbytes);
    if (e == cudaSuccess) e = cudaMemset(acc[i], 0, abytes);
    if (e != cudaSuccess)
    {
      /* the plan keeps the rings it had; nothing half-built stays behind */
      for (void* q : hist) if (q) cudaFree(q);
      for (void* q : acc) if (q) cudaFree(q);
      plan_fail(p, (int)e, "state rings", __FILE__, __LINE__);
      return false;
    }
  }
  if (!p->history.empty())
  {
    CU_TRY(p, cudaMemcpy(hist[0], p->history[p->state_sel], hbytes, cudaMemcpyDeviceToDevice));
    CU_TRY(p, cudaMemcpy(acc[0], p->acc_state[p->state_sel], abytes, cudaMemcpyDeviceToDevice));
    for (void* q : p->history) cudaFree(q);
    for (void* q : p->acc_state) cudaFree(q);
  }
  p->history.swap(hist);
  p->acc_state.swap(acc);
  p->state_sel = 0;
  for (Plan::Slot& s : p->slots)
    for (Buffer* b : { &s.prefix, &s.chain_totals, &s.flags })
      if (b->ptr) cudaFree(b->ptr);
  p->slots.assign(depth + 2, Plan::Slot());      // streaming slots, the serial slot, the tail slot of split calls
  if (p->control) cudaFree(p->control);
  p->control = nullptr;
  CU_TRY(p, cudaMalloc(&p->control, (2 + 4 * ((size_t)depth + 2)) * sizeof(unsigned)));
  CU_TRY(p, cudaMemset(p->control, 0, (2 + 4 * ((size_t)depth + 2)) * sizeof(unsigned)));
  p->stream_depth = depth;
  p->calls_issued = 0;
  p->prev_slot = 0;
  return true;
}

template <typename T, typename F>
bool plan_build(Plan* p)
{
  const size_t m = p->m, cells = p->cells;
  std::vector<cx<F>> tw, tws;
  make_tables<F>(m, p->latency, tw, tws);

  std::vector<cx<F>> tw_ext(cells), p0(cells);
  for (size_t k = 0; k < m; ++k)
  {
    tw_ext[k + 2] = tw[k];
    p0[k + 2].r = (F)1; p0[k + 2].i = (F)0;
  }
  for (int q = 0; q < 4; ++q)
  {
    const int c = p->mirrors.cell[q], s = p->mirrors.src[q];
    if (s < 0)
    {
      tw_ext[c].r = tw_ext[c].i = (F)0;
      p0[c].r = p0[c].i = (F)0;
    }
    else
    {
      tw_ext[c] = tw[s];
      if (p->mirrors.conj[q]) tw_ext[c].i = -tw_ext[c].i;
      p0[c].r = (F)1; p0[c].i = (F)0;
    }
  }

  const bool table = (type_id<F>::value == kF32);      // float: recurrence table; double: roots of unity (PhaseSource)
  if (table)
  {
    const size_t budget = env_size("SDFT_B200_F0_BUDGET_MB", 192) << 20;
    p->f0_stride = f0_stride_for(m, sizeof(cx<F>), budget);
    p->f0_rows = (2 * m + p->f0_stride - 1) / p->f0_stride;
    size_t free_b = 0, total_b = 0;
    CU_TRY(p, cudaMemGetInfo(&free_b, &total_b));
    if (p->f0_rows * cells * sizeof(cx<F>) > free_b / 4)
    {
      plan_fail(p, SDFT_B200_ERR_ARG, "sdft_alloc: dftsize too large for the float phase table on this device "
                                      "(float frequency domain needs the reference's sequential recurrence: "
                                      "m^2/64 bytes once the table stride has reached the longest chunk)", __FILE__, __LINE__);
      return false;
    }
  }
  else
  {
    p->f0_stride = kF0Stride;
    p->f0_rows = 1;
  }
  CU_TRY(p, cudaMalloc(&p->tw_ext, cells * sizeof(cx<F>)));
  CU_TRY(p, cudaMalloc(&p->tws, m * sizeof(cx<F>)));
  CU_TRY(p, cudaMalloc(&p->f0, p->f0_rows * cells * sizeof(cx<F>)));
  p->table_bytes = (cells + m + p->f0_rows * cells) * sizeof(cx<F>);
  if (!table)
  {
    /* E[j] = exp(-2 pi i j / 2m), evaluated in long double and rounded once */
    std::vector<cx<F>> roots(2 * m);
    const long double step = -2.0L * acosl(-1.0L) / (long double)(2 * m);
    for (size_t j = 0; j < 2 * m; ++j)
    {
      roots[j].r = (F)cosl(step * (long double)j);
      roots[j].i = (F)sinl(step * (long double)j);
    }
    CU_TRY(p, cudaMalloc(&p->roots, 2 * m * sizeof(cx<F>)));
    CU_TRY(p, cudaMemcpy(p->roots, roots.data(), 2 * m * sizeof(cx<F>), cudaMemcpyHostToDevice));
    p->table_bytes += 2 * m * sizeof(cx<F>);
  }
  if (!plan_rings<T, F>(p, 1)) return false;
  CU_TRY(p, cudaMalloc(&p->phase_scratch, cells * sizeof(cx<F>)));

  CU_TRY(p, cudaMemcpyAsync(p->tw_ext, tw_ext.data(), cells * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
  CU_TRY(p, cudaMemcpyAsync(p->tws, tws.data(), m * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
  /* stage P0 (1, or 0 for always-zero mirror cells), expand it into the table */
  CU_TRY(p, cudaMemcpyAsync(p->phase_scratch, p0.data(), cells * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
  if (table)
  {
    const unsigned threads = 128;
    phase_table_kernel<F><<<(unsigned)((cells + threads - 1) / threads), threads, 0, p->stream>>>(
        (const cx<F>*)p->tw_ext, (const cx<F>*)p->phase_scratch, (cx<F>*)p->f0, (unsigned)cells, (unsigned)(2 * m), p->f0_stride);
    p->launches++;
    CU_TRY(p, cudaGetLastError());
  }
  else
  {
    CU_TRY(p, cudaMemcpyAsync(p->f0, p0.data(), cells * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
  }
  CU_TRY(p, cudaStreamSynchronize(p->stream));   // host vectors go out of scope
  return true;
}

template <typename T, typename F>
bool plan_reset(Plan* p)
{
  const size_t m = p->m, cells = p->cells, ch = p->channels;
  p->cursor = 0;
  /* stream-ordered: the memsets run after every call queued so far has completed (calls complete in order) */
  p->state_sel = 0;
  CU_TRY(p, cudaMemsetAsync(p->history[0], 0, ch * 2 * m * sizeof(T), p->stream));
  for (void* a : p->acc_state) CU_TRY(p, cudaMemsetAsync(a, 0, ch * cells * sizeof(cx<F>), p->stream));
  return true;
}

void plan_destroy(Plan* p)
{
  if (!p) return;
  DeviceGuard on_device(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
  for (void* q : p->history) cudaFree(q);
  for (void* q : p->acc_state) cudaFree(q);
  for (Plan::Slot& s : p->slots)
    for (Buffer* b : { &s.prefix, &s.chain_totals, &s.flags })
      if (b->ptr) cudaFree(b->ptr);
  void* ptrs[] = { p->tw_ext, p->tws, p->f0, p->roots, p->phase_scratch, p->control,
                   p->samples.ptr, p->synth_out.ptr, p->row_ptrs.ptr, p->part.ptr, p->weights.ptr, p->syn_ab.ptr, p->trace.ptr, p->tile[0].ptr, p->tile[1].ptr };
  for (void* q : ptrs)
    if (q) cudaFree(q);
  for (int w = 0; w < 2; ++w)
    for (cudaEvent_t e : p->prof_events[w]) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
  {
    if (p->stage[i]) cudaFreeHost(p->stage[i]);
    if (i == 0 && p->mailbox) cudaFreeHost(p->mailbox);
    if (p->stage_done[i]) cudaEventDestroy(p->stage_done[i]);
    if (p->tile_ready[i]) cudaEventDestroy(p->tile_ready[i]);
    if (p->tile_free[i]) cudaEventDestroy(p->tile_free[i]);
  }
  if (p->samples_in) cudaEventDestroy(p->samples_in);
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  cudaGetLastError();
  delete p;
}

template <typename T, typename F>
Plan* plan_create(size_t m, int window, double latency, size_t channels)
{
  g_alloc_error = 0;
  g_alloc_errmsg[0] = 0;
  if (m == 0 || channels == 0 || channels > 65535 || m > (1u << 30) || window < 0 || window > 3)
  {
    plan_fail(nullptr, SDFT_B200_ERR_ARG, "sdft_alloc: bad dftsize/window/channels", __FILE__, __LINE__);
    return nullptr;
  }
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0)
  {
    cudaGetLastError();
    plan_fail(nullptr, SDFT_B200_ERR_NODEVICE, "no CUDA device (libsdft_b200 has no CPU fallback)", __FILE__, __LINE__);
    return nullptr;
  }
  Plan* p = new (std::nothrow) Plan();
  if (!p) return nullptr;
  p->td = type_id<T>::value;
  p->fd = type_id<F>::value;
  p->m = m;
  p->cells = m + 4;
  p->window = window;
  p->latency = latency;
  p->channels = channels;
  p->mirrors = make_mirrors(m);
  p->tile_bytes = env_size("SDFT_B200_TILE_MB", 128) << 20;
  p->forced_chunk = env_size("SDFT_B200_CHUNK", 0);
  p->forced_warps = (unsigned)env_size("SDFT_B200_WARPS", 0);
  {
    const char* pg = getenv("SDFT_B200_PAGEABLE");
    p->driver_pageable = pg && !strcmp(pg, "driver");
    p->pdl = env_size("SDFT_B200_PDL", 1) != 0;
    p->no_split = env_size("SDFT_B200_NO_SPLIT", 0) != 0;
    p->mailbox_on = env_size("SDFT_B200_MAILBOX", 1) != 0;
    const char* ge = getenv("SDFT_B200_GEO");
    if (ge && !strcmp(ge, "wide")) p->forced_geo = GEO_WIDE;
    if (ge && !strcmp(ge, "narrow")) p->forced_geo = GEO_NARROW;
  }
  if (p->forced_warps > (unsigned)kScanWarps) p->forced_warps = kScanWarps;
  {
    /* double frequency domain: fast (demodulated replay) unless SDFT_B200_F64=modulated.
     * float: the reference's modulated scheme; the replay keeps every rounding of the reference (rows are
     * bit-exact within a chunk); the chunk totals that feed the carries are summed in double on the FP64
     * pipe unless SDFT_B200_F32=strict asks for the float recurrence there too */
    if (type_id<F>::value == kF64)
    {
      const char* md = getenv("SDFT_B200_F64");
      p->mode = (md && !strcmp(md, "modulated")) ? MODE_MODULATED : MODE_FAST;
    }
    else
    {
      const char* md = getenv("SDFT_B200_F32");
      p->mode = (md && !strcmp(md, "strict")) ? MODE_MODULATED : MODE_FAST;
    }
    p->prescale = (p->mode == MODE_FAST && type_id<F>::value == kF64) ? (double)make_window_const<double>(m, window).pre : 1.0;
  }

  bool ok = true;
  const long dev_env = (long)env_size("SDFT_B200_DEVICE", (size_t)-1);
  cudaError_t e = cudaGetDevice(&p->device);
  if (e == cudaSuccess && dev_env >= 0)
  {
    if (dev_env < count) p->device = (int)dev_env;
    else e = cudaErrorInvalidDevice;
  }
  DeviceGuard on_device(p->device);     // SDFT_B200_DEVICE places the plan, it does not move the caller
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->samples_in, cudaEventDisableTiming);
  for (int i = 0; i < 2 && e == cudaSuccess; ++i)
  {
    e = cudaEventCreateWithFlags(&p->tile_ready[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->stage_done[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&p->tile_free[i], cudaEventDisableTiming);
  }
  if (e != cudaSuccess)
  {
    plan_fail(p, (int)e, "plan_create: stream/event setup", __FILE__, __LINE__);
    ok = false;
  }
  p->stream = p->own_stream;
  ok = ok && plan_build<T, F>(p) && plan_reset<T, F>(p);
  if (ok && cudaStreamSynchronize(p->stream) != cudaSuccess) ok = false;
  if (!ok)
  {
    g_alloc_error = p->status ? p->status : (int)cudaGetLastError();
    snprintf(g_alloc_errmsg, sizeof(g_alloc_errmsg), "%s", p->errmsg);
    plan_destroy(p);
    return nullptr;
  }
  return p;
}

}  // namespace
