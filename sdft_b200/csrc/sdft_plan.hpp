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
  size_t tile_bytes = 0;
  unsigned long long launches = 0;

  MirrorMap mirrors;
  int mode = 0;                  // MODE_MODULATED / MODE_FAST (double frequency domain only)
  double prescale = 1.0;         // factor folded into the deltas in fast mode (acc_state is scaled by it)
  void* tw_ext = nullptr;
  void* tws = nullptr;
  void* f0 = nullptr;
  size_t f0_rows = 0;

  void* history[2] = { nullptr, nullptr };
  int hist_sel = 0;
  void* acc_state[2] = { nullptr, nullptr };     // ping-pong, same reason (neighbouring groups share halo cells)
  int acc_sel = 0;
  void* phase_scratch = nullptr;   // cells complex values, introspection only

  Buffer samples, synth_out, tile[2], part, weights;
  Buffer syn_ab;                 // fused synthesis: per-bin (A, B) weights with the window folded in (make_synth_weights)
  bool syn_ab_ready = false, syn_ab_unit = false;
  Buffer trace;                  // -DSDFT_B200_TRACE builds: per-CTA phase stamps of the last scan launch
  size_t trace_items = 0;
  void* stage[2] = { nullptr, nullptr };   // pinned host staging for PAGEABLE caller buffers (grow-only)
  size_t stage_bytes[2] = { 0, 0 };
  cudaEvent_t stage_done[2] = { nullptr, nullptr };
  Buffer prefix, chain_totals, flags;   // chained scan: inclusive prefixes, chunk totals, epoch-stamped flags
  unsigned* control = nullptr;   // [0] work ticket, [1] spin-wait timeout flag
  unsigned epoch = 0;

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

/* -------- plan construction -------- */
template <typename T, typename F>
bool plan_build(Plan* p)
{
  const size_t m = p->m, cells = p->cells, ch = p->channels;
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

  p->f0_rows = (2 * m + kF0Stride - 1) / kF0Stride;
  CU_TRY(p, cudaMalloc(&p->tw_ext, cells * sizeof(cx<F>)));
  CU_TRY(p, cudaMalloc(&p->tws, m * sizeof(cx<F>)));
  CU_TRY(p, cudaMalloc(&p->f0, p->f0_rows * cells * sizeof(cx<F>)));
  CU_TRY(p, cudaMalloc(&p->history[0], ch * 2 * m * sizeof(T)));
  CU_TRY(p, cudaMalloc(&p->history[1], ch * 2 * m * sizeof(T)));
  CU_TRY(p, cudaMalloc(&p->acc_state[0], ch * cells * sizeof(cx<F>)));
  CU_TRY(p, cudaMalloc(&p->acc_state[1], ch * cells * sizeof(cx<F>)));
  CU_TRY(p, cudaMalloc(&p->control, 2 * sizeof(unsigned)));
  CU_TRY(p, cudaMemsetAsync(p->control, 0, 2 * sizeof(unsigned), p->stream));
  CU_TRY(p, cudaMalloc(&p->phase_scratch, cells * sizeof(cx<F>)));

  CU_TRY(p, cudaMemcpyAsync(p->tw_ext, tw_ext.data(), cells * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
  CU_TRY(p, cudaMemcpyAsync(p->tws, tws.data(), m * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
  /* stage P0 (1, or 0 for always-zero mirror cells), expand it into the table */
  CU_TRY(p, cudaMemcpyAsync(p->phase_scratch, p0.data(), cells * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
  const unsigned threads = 128;
  phase_table_kernel<F><<<(unsigned)((cells + threads - 1) / threads), threads, 0, p->stream>>>(
      (const cx<F>*)p->tw_ext, (const cx<F>*)p->phase_scratch, (cx<F>*)p->f0, (unsigned)cells, (unsigned)(2 * m));
  p->launches++;
  CU_TRY(p, cudaGetLastError());
  CU_TRY(p, cudaStreamSynchronize(p->stream));   // host vectors go out of scope
  return true;
}

template <typename T, typename F>
bool plan_reset(Plan* p)
{
  const size_t m = p->m, cells = p->cells, ch = p->channels;
  p->cursor = 0;
  p->hist_sel = 0;
  p->acc_sel = 0;
  CU_TRY(p, cudaMemsetAsync(p->history[0], 0, ch * 2 * m * sizeof(T), p->stream));
  CU_TRY(p, cudaMemsetAsync(p->acc_state[0], 0, ch * cells * sizeof(cx<F>), p->stream));
  CU_TRY(p, cudaMemsetAsync(p->acc_state[1], 0, ch * cells * sizeof(cx<F>), p->stream));
  return true;
}

void plan_destroy(Plan* p)
{
  if (!p) return;
  DeviceGuard on_device(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
  void* ptrs[] = { p->tw_ext, p->tws, p->f0, p->history[0], p->history[1], p->acc_state[0], p->acc_state[1], p->phase_scratch,
                   p->prefix.ptr, p->chain_totals.ptr, p->flags.ptr, p->control,
                   p->samples.ptr, p->synth_out.ptr, p->part.ptr, p->weights.ptr, p->syn_ab.ptr, p->trace.ptr, p->tile[0].ptr, p->tile[1].ptr };
  for (void* q : ptrs)
    if (q) cudaFree(q);
  for (int w = 0; w < 2; ++w)
    for (cudaEvent_t e : p->prof_events[w]) cudaEventDestroy(e);
  for (int i = 0; i < 2; ++i)
  {
    if (p->stage[i]) cudaFreeHost(p->stage[i]);
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
