/*
 * sdft_b200.cu -- host side of libsdft_b200.so: device-resident plan, launch logic and the C-ABI
 * declared in include/sdft_b200.h.  The kernels are in sdft_kernels.cuh.
 *
 * What lives where (reference: struct sdft_plan, c/src/sdft/sdft.h:137-182):
 *   tables   tw_ext[m+4], tws[m], F0[ceil(2m/32)][m+4]         device, written once per plan
 *   state    history[2][2m] and acc_state[2][m+4] (ping-pong) per channel; the cursor lives on the host and the
 *            modulation phase is a pure function of it (table row + <32 rotations), so it is not stored
 *   scratch  samples, deltas, chunk totals/prefixes/flags of the chained scan, row tiles   device, grow-only
 *
 * No CPU fallback: every entry point either runs the CUDA path or records an error on the plan.
 */
#include "../../include/sdft_b200.h"
#include "sdft_kernels.cuh"
#include "sdft_host_copy.hpp"

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

/* ------------------------------------------------------------------------------------------------
 * plan
 * ---------------------------------------------------------------------------------------------- */
enum TypeId { kF32 = 0, kF64 = 1 };

/* mirror cells (sdft.h:589-595): source bin (or -1 = always zero) and whether the copy is conjugated;
 * only used on the host to lay out the extended twiddle table */
struct MirrorMap
{
  int cell[4];
  int src[4];
  int conj[4];
};

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
  cudaEvent_t tile_ready[2] = { nullptr, nullptr };
  cudaEvent_t tile_free[2] = { nullptr, nullptr };

  size_t cursor = 0;
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
 * host-side trigonometry with the reference's expression order and types (sdft.h:439-446).
 * Computed with the host libm so that float tables are bit-identical to the reference's
 * (SURVEY.md fact 5); the device never evaluates sin/cos.
 * ---------------------------------------------------------------------------------------------- */
inline float t_cos(float x) { return ::cosf(x); }
inline float t_sin(float x) { return ::sinf(x); }
inline float t_acos(float x) { return ::acosf(x); }
inline double t_cos(double x) { return ::cos(x); }
inline double t_sin(double x) { return ::sin(x); }
inline double t_acos(double x) { return ::acos(x); }

template <typename F>
void make_tables(size_t m, double latency, std::vector<cx<F>>& tw, std::vector<cx<F>>& tws)
{
  tw.resize(m);
  tws.resize(m);
  const F omega = (F)(-2) * t_acos((F)(-1)) / (F)(m * 2);
  const F wsyn = (F)(+2) / ((F)(1) - t_cos((F)((omega * (F)m) * latency)));
  for (size_t k = 0; k < m; ++k)
  {
    const F a = omega * (F)k;
    tw[k].r = (F)(1) * t_cos(a);
    tw[k].i = (F)(1) * t_sin(a);
    const F s = (F)(((omega * (F)k) * (F)m) * latency);   // trailing product in double, then narrowed
    tws[k].r = wsyn * t_cos(s);
    tws[k].i = wsyn * t_sin(s);
  }
}

/* mirror cells resolved from the assignment order of sdft.h:589-595 */
MirrorMap make_mirrors(size_t m)
{
  MirrorMap mm;
  mm.cell[0] = 0; mm.cell[1] = 1; mm.cell[2] = (int)m + 2; mm.cell[3] = (int)m + 3;
  if (m >= 3)
  {
    mm.src[0] = 2; mm.conj[0] = 1;
    mm.src[1] = 1; mm.conj[1] = 1;
    mm.src[2] = (int)m - 2; mm.conj[2] = 1;
    mm.src[3] = (int)m - 3; mm.conj[3] = 1;
  }
  else if (m == 2)
  {
    /* aux[1]=conj(bin1); aux[4]=conj(bin0); aux[0]=conj(aux[4])=bin0; aux[5]=conj(aux[1])=bin1 */
    mm.src[0] = 0; mm.conj[0] = 0;
    mm.src[1] = 1; mm.conj[1] = 1;
    mm.src[2] = 0; mm.conj[2] = 1;
    mm.src[3] = 1; mm.conj[3] = 0;
  }
  else
  {
    /* m == 1: each mirror cell only ever copies itself through its partner and stays zero */
    for (int q = 0; q < 4; ++q) { mm.src[q] = -1; mm.conj[q] = 0; }
  }
  return mm;
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
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  if (p->copy_stream) cudaStreamSynchronize(p->copy_stream);
  void* ptrs[] = { p->tw_ext, p->tws, p->f0, p->history[0], p->history[1], p->acc_state[0], p->acc_state[1], p->phase_scratch,
                   p->prefix.ptr, p->chain_totals.ptr, p->flags.ptr, p->control,
                   p->samples.ptr, p->synth_out.ptr, p->part.ptr, p->weights.ptr, p->trace.ptr, p->tile[0].ptr, p->tile[1].ptr };
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
  cudaError_t e = cudaSuccess;
  if (dev_env >= 0) e = cudaSetDevice((int)dev_env);
  if (e == cudaSuccess) e = cudaGetDevice(&p->device);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&p->copy_stream, cudaStreamNonBlocking);
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

/* warp-wide groups of bins the plan's m bins are cut into, for either warp geometry */
unsigned groups_for(const Plan* p, int geo = GEO_WIDE)
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
  if (p->window == 0) halo = 0;
  const unsigned span = wc - 2 * halo;
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

/* 32-byte group stores need rows that start on a 32-byte boundary */
template <typename F>
bool can_vectorize(size_t m, const void* out, size_t out_stride)
{
  const size_t g = Geo<F, GEO_WIDE>::GROUP;
  return (m % g == 0) && (((uintptr_t)out) % 32 == 0) && (out_stride % g == 0);
}

unsigned choose_chunk(const Plan* p, size_t n, int geo)
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
     * has fewer than 16 chunks */
    unsigned c = (u < 16384.0) ? 32u : ((u < 30.0e3) ? 64u : 128u);
    while (c > 32u && (size_t)c * 16 > n) c >>= 1;
    return c;
  }
  if (u < 16384.0) return 32;
  if (u < 100.0e3) return 64;
  if (u < 4.0e6) return 128;
  if (u < 16.0e6) return 256;
  return kAutoChunk;
}

template <typename F, int EMIT, int GEO>
void launch_chain_geo(Plan* p, const ChainArgs<F>& a, bool vec, unsigned warps)
{
  const dim3 grid(a.total_blocks);
  const size_t smem = scan_smem_bytes<F, GEO>(warps, a.sched.chunk) + (size_t)a.stage_rows * Geo<F, GEO>::WC * sizeof(cx<F>);
  constexpr int kDefaultMode = (int)MODE_FAST;
  /* programmatic dependent launch: the CTAs of this call may become resident while the previous kernel
   * of the stream drains; they wait at the top of the kernel (griddepcontrol.wait) until that kernel
   * has completed and flushed, so nothing else about the ordering changes.  Hides the launch latency
   * between back-to-back calls (streaming). */
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
  if (GEO == GEO_NARROW || p->mode == kDefaultMode)
  {
    /* the narrow geometry exists for the default mode only (choose_geo) */
    switch (p->window)
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
    switch (p->window)
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

/* warps (= consecutive chunks) per scan/emit CTA */
unsigned scan_warps_for(const Plan* p, unsigned chunk, unsigned nchunks)
{
  /* 4 is the measured optimum: wider CTAs shorten the global chain further but pile their stores onto
   * one SM, narrower ones lengthen the chain */
  unsigned w = p->forced_warps ? p->forced_warps : 4u;
  if (w > (unsigned)kSmemSamples / chunk) w = (unsigned)kSmemSamples / chunk;
  if (w > (unsigned)kScanWarps) w = kScanWarps;
  if (w > nchunks) w = nchunks;
  if (w < 1) w = 1;
  return w;
}

/* production path: the single-pass chained scan/emit kernel (deltas and history are its prologue) */
template <typename T, typename F>
bool analysis_chained(Plan* p, size_t n, const T* x, size_t x_stride, cx<F>* out, size_t out_stride, F* part = nullptr,
                      const cx<F>* weights = nullptr)
{
  const unsigned m = (unsigned)p->m;
  const unsigned ch = (unsigned)p->channels;
  const int geo = choose_geo(p, n);
  const unsigned chunk = choose_chunk(p, n, geo);
  const Schedule sched = make_schedule(p->cursor, n, m, chunk);
  const unsigned groups = groups_for(p, geo);
  const size_t wc = (geo == GEO_NARROW) ? (size_t)Geo<F, GEO_NARROW>::WC : (size_t)Geo<F, GEO_WIDE>::WC;
  const unsigned warps = scan_warps_for(p, chunk, sched.nchunks);
  const unsigned nblocks = (sched.nchunks + warps - 1) / warps;
  const size_t items = (size_t)ch * nblocks * groups;
  if (items >= (1ull << 31))
  {
    plan_fail(p, SDFT_B200_ERR_ARG, "analysis: call too large for one launch", __FILE__, __LINE__);
    return false;
  }

  if (!reserve(p, p->prefix, items * wc * sizeof(cx<F>))) return false;
  if (!reserve(p, p->chain_totals, items * wc * sizeof(cx<F>))) return false;
  const size_t flags_before = p->flags.bytes;
  if (!reserve(p, p->flags, items * sizeof(unsigned))) return false;
  if (p->flags.bytes != flags_before || p->epoch >= 0x7ffffff0u)
  {
    CU_TRY(p, cudaMemsetAsync(p->flags.ptr, 0, p->flags.bytes, p->stream));
    p->epoch = 0;
  }
  p->epoch++;

  ChainArgs<F> a;
  a.sched = sched;
  a.samples = x;
  a.sample_stride = x_stride;
  a.hist_old = p->history[p->hist_sel];
  a.hist_new = p->history[p->hist_sel ^ 1];
  a.td_double = (type_id<T>::value == kF64) ? 1 : 0;
  a.scale = (F)p->prescale;
  a.tw_ext = (const cx<F>*)p->tw_ext;
  a.f0 = (const cx<F>*)p->f0;
  a.acc_in = (const cx<F>*)p->acc_state[p->acc_sel];
  a.acc_out = (cx<F>*)p->acc_state[p->acc_sel ^ 1];
  a.prefix = (cx<F>*)p->prefix.ptr;
  a.totals = (cx<F>*)p->chain_totals.ptr;
  a.flags = (unsigned*)p->flags.ptr;
  a.control = p->control;
  a.epoch = p->epoch;
  a.total_blocks = (unsigned)items;
  a.nblocks = nblocks;
  a.channels = ch;
  a.m = m;
  a.cells = (unsigned)p->cells;
  a.out = out;
  a.out_channel_stride = out_stride;
  a.tws = weights ? weights : (const cx<F>*)p->tws;
  a.part = part;
  a.groups = groups;
  a.stage_rows = (geo == GEO_NARROW) ? scan_stage_rows<F, GEO_NARROW>(warps, chunk) : scan_stage_rows<F, GEO_WIDE>(warps, chunk);
  a.win = make_window_const<F>(p->m, p->window);   // sdft.h:422, :371
  a.trace = nullptr;
#if defined(SDFT_B200_TRACE)
  if (reserve(p, p->trace, items * 8 * sizeof(unsigned long long)))
  {
    a.trace = (unsigned long long*)p->trace.ptr;
    p->trace_items = items;
  }
#endif
  if (part)
  {
    prof_mark(p, 0);
    if (p->latency == 1 && !weights) launch_chain<F, EMIT_SYNTH_UNIT>(p, a, false, warps, geo);   // exact compare, sdft.h:639
    else launch_chain<F, EMIT_SYNTH>(p, a, false, warps, geo);
    prof_mark(p, 0);
  }
  else if (out)
  {
    const bool vec = can_vectorize<F>(m, out, out_stride);
    prof_mark(p, 0);
    launch_chain<F, EMIT_ROWS>(p, a, vec, warps, geo);
    prof_mark(p, 0);
  }
  else
  {
    launch_chain<F, EMIT_NONE>(p, a, false, warps, geo);
  }
  CU_TRY(p, cudaGetLastError());
  p->hist_sel ^= 1;
  p->cursor = (size_t)((p->cursor + n) % (2 * (size_t)m));
  p->acc_sel ^= 1;
  return true;
}

/* analysis over n samples per channel, everything on the device.
 * x: (channels, x_stride) samples; out: (channels, out_stride) complex rows or nullptr (state only). */
template <typename T, typename F>
bool analysis_device(Plan* p, size_t n, const T* x, size_t x_stride, cx<F>* out, size_t out_stride)
{
  if (n == 0) return true;
  return analysis_chained<T, F>(p, n, x, x_stride, out, out_stride);
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
                                                                       y_stride, n, (unsigned)p->m);
  else
    synth_kernel<T, F, false><<<grid, kSynthWarps * 32, 0, p->stream>>>(dfts, dft_stride, (const cx<F>*)p->tws, y,
                                                                        y_stride, n, (unsigned)p->m);
  prof_mark(p, 1);
  p->launches++;
  CU_TRY(p, cudaGetLastError());
  return true;
}

/* ------------------------------------------------------------------------------------------------
 * host/device pointer plumbing
 * ---------------------------------------------------------------------------------------------- */
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
  if (cudaMemcpyAsync(p->samples.ptr, samples, bytes, cudaMemcpyHostToDevice, p->stream) != cudaSuccess)
  {
    plan_fail(p, (int)cudaGetLastError(), "H2D samples", __FILE__, __LINE__);
    *ok = false;
    return nullptr;
  }
  return (const T*)p->samples.ptr;
}

template <typename T, typename F>
bool do_sdft(Plan* p, size_t n, const T* samples, cx<F>* dfts)
{
  if (n == 0) return true;
  CU_TRY(p, cudaSetDevice(p->device));
  bool ok = true;
  const T* x = stage_samples<T>(p, n, samples, &ok);
  if (!ok) return false;
  const size_t m = p->m, ch = p->channels;

  if (classify(dfts) == kDevice)
  {
    return analysis_device<T, F>(p, n, x, n, dfts, n * m);
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
  return true;
}

template <typename T, typename F>
bool do_advance(Plan* p, size_t n, const T* samples)
{
  if (n == 0) return true;
  CU_TRY(p, cudaSetDevice(p->device));
  bool ok = true;
  const bool host = classify(samples) != kDevice;
  const T* x = stage_samples<T>(p, n, samples, &ok);
  if (!ok) return false;
  if (!analysis_device<T, F>(p, n, x, n, (cx<F>*)nullptr, 0)) return false;
  if (host) CU_TRY(p, cudaStreamSynchronize(p->stream));
  return true;
}

template <typename T, typename F>
bool do_isdft(Plan* p, size_t n, const cx<F>* dfts, T* samples)
{
  if (n == 0) return true;
  CU_TRY(p, cudaSetDevice(p->device));
  const size_t m = p->m, ch = p->channels;
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
  return true;
}

/* row-pointer variants (sdft.h:622-628, 681-687): rows may be host or device pointers */
template <typename T, typename F>
bool do_sdft_nd(Plan* p, size_t n, const T* samples, cx<F>** rows_out)
{
  if (n == 0) return true;
  if (p->channels != 1) { plan_fail(p, SDFT_B200_ERR_ARG, "sdft_nd on a batch plan", __FILE__, __LINE__); return false; }
  CU_TRY(p, cudaSetDevice(p->device));
  bool ok = true;
  const T* x = stage_samples<T>(p, n, samples, &ok);
  if (!ok) return false;
  const size_t m = p->m, row_bytes = m * sizeof(cx<F>);
  const size_t rows = tile_rows(p, n, row_bytes);
  if (!reserve(p, p->tile[0], rows * row_bytes)) return false;
  for (size_t t0 = 0; t0 < n; t0 += rows)
  {
    const size_t len = (t0 + rows <= n) ? rows : n - t0;
    if (!analysis_device<T, F>(p, len, x + t0, n, (cx<F>*)p->tile[0].ptr, len * m)) return false;
    for (size_t i = 0; i < len; ++i)
    {
      CU_TRY(p, cudaMemcpyAsync(rows_out[t0 + i], (cx<F>*)p->tile[0].ptr + i * m, row_bytes, cudaMemcpyDefault, p->stream));
    }
    CU_TRY(p, cudaStreamSynchronize(p->stream));
  }
  return true;
}

template <typename T, typename F>
bool do_isdft_nd(Plan* p, size_t n, const cx<F>** rows_in, T* samples)
{
  if (n == 0) return true;
  if (p->channels != 1) { plan_fail(p, SDFT_B200_ERR_ARG, "isdft_nd on a batch plan", __FILE__, __LINE__); return false; }
  CU_TRY(p, cudaSetDevice(p->device));
  const size_t m = p->m, row_bytes = m * sizeof(cx<F>);
  const size_t rows = tile_rows(p, n, row_bytes);
  if (!reserve(p, p->tile[0], rows * row_bytes)) return false;
  const bool out_dev = classify(samples) == kDevice;
  T* y = samples;
  if (!out_dev)
  {
    if (!reserve(p, p->synth_out, n * sizeof(T))) return false;
    y = (T*)p->synth_out.ptr;
  }
  for (size_t t0 = 0; t0 < n; t0 += rows)
  {
    const size_t len = (t0 + rows <= n) ? rows : n - t0;
    for (size_t i = 0; i < len; ++i)
    {
      CU_TRY(p, cudaMemcpyAsync((cx<F>*)p->tile[0].ptr + i * m, rows_in[t0 + i], row_bytes, cudaMemcpyDefault, p->stream));
    }
    if (!synthesis_device<T, F>(p, len, (const cx<F>*)p->tile[0].ptr, len * m, y + t0, n)) return false;
    CU_TRY(p, cudaStreamSynchronize(p->stream));
  }
  if (!out_dev)
  {
    CU_TRY(p, cudaMemcpyAsync(samples, y, n * sizeof(T), cudaMemcpyDeviceToHost, p->stream));
    CU_TRY(p, cudaStreamSynchronize(p->stream));
  }
  return true;
}

/* analysis -> synthesis in ONE kernel: the rows never exist in memory.  Every warp weighs and reduces
 * its bins per time step (SynthLane), per-group partial sums go to a scratch buffer and a small second
 * kernel adds the groups in order.  Long calls are cut into pieces that bound the scratch. */
template <typename T, typename F>
bool do_roundtrip(Plan* p, size_t n, const T* in, T* out, const cx<F>* gains = nullptr)
{
  if (n == 0) return true;
  CU_TRY(p, cudaSetDevice(p->device));
  const cx<F>* weights = nullptr;
  if (gains)
  {
    /* spectral processing between analysis and synthesis: every row is multiplied bin by bin with
     * `gains` before sdft_isdft sees it, i.e. the synthesis weights become gains[k] * tws[k]
     * (gains[k] * (-1)^k for latency 1, sdft.h:639-652) */
    const size_t m = p->m;
    std::vector<cx<F>> g(m), w(m);
    if (classify(gains) == kDevice) CU_TRY(p, cudaMemcpy(g.data(), gains, m * sizeof(cx<F>), cudaMemcpyDeviceToHost));
    else memcpy(g.data(), gains, m * sizeof(cx<F>));
    std::vector<cx<F>> tw, tws;
    make_tables<F>(m, p->latency, tw, tws);
    for (size_t k = 0; k < m; ++k)
    {
      cx<F> t = tws[k];
      if (p->latency == 1) { t.r = (k & 1) ? (F)(-1) : (F)(1); t.i = (F)0; }
      w[k].r = g[k].r * t.r - g[k].i * t.i;
      w[k].i = g[k].r * t.i + g[k].i * t.r;
    }
    if (!reserve(p, p->weights, m * sizeof(cx<F>))) return false;
    CU_TRY(p, cudaMemcpyAsync(p->weights.ptr, w.data(), m * sizeof(cx<F>), cudaMemcpyHostToDevice, p->stream));
    CU_TRY(p, cudaStreamSynchronize(p->stream));   // w goes out of scope
    weights = (const cx<F>*)p->weights.ptr;
  }
  bool ok = true;
  const T* x = stage_samples<T>(p, n, in, &ok);
  if (!ok) return false;
  const size_t ch = p->channels;
  const unsigned max_groups = groups_for(p, GEO_NARROW);    // either geometry may be chosen per piece
  const bool out_dev = classify(out) == kDevice;
  T* y = out;
  if (!out_dev)
  {
    if (!reserve(p, p->synth_out, ch * n * sizeof(T))) return false;
    y = (T*)p->synth_out.ptr;
  }
  size_t piece = env_size("SDFT_B200_ROUNDTRIP_PIECE", (size_t)1 << 22);
  if (piece > n) piece = n;
  if (!reserve(p, p->part, ch * max_groups * piece * sizeof(F))) return false;
  for (size_t t0 = 0; t0 < n; t0 += piece)
  {
    const size_t len = (t0 + piece <= n) ? piece : n - t0;
    const unsigned groups = groups_for(p, choose_geo(p, len));   // what analysis_chained will use for this piece
    if (!analysis_chained<T, F>(p, len, x + t0, n, (cx<F>*)nullptr, 0, (F*)p->part.ptr, weights)) return false;
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
  }
  return true;
}

/* SDFT.convolve of the reference's Python class (python/src/sdft/sdft.py:146-203): window(rows) / m */
template <typename F>
bool do_convolve(Plan* p, size_t n, const cx<F>* in, cx<F>* out)
{
  if (n == 0) return true;
  CU_TRY(p, cudaSetDevice(p->device));
  const size_t m = p->m, ch = p->channels;
  const int need = (p->window == 3) ? 3 : ((p->window == 0) ? 1 : 2);
  if ((int)m < need)
  {
    plan_fail(p, SDFT_B200_ERR_ARG, "convolve: dftsize too small for this window", __FILE__, __LINE__);
    return false;
  }
  const size_t bytes = ch * n * m * sizeof(cx<F>);
  const bool in_dev = classify(in) == kDevice, out_dev = classify(out) == kDevice;
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
    if (typed<TD, FD>(p, "sdft_reset: plan type mismatch")) plan_reset<TD, FD>(p);                              \
  }                                                                                                             \
  extern "C" size_t sdft_b200_##SFX##_size(const sdft_b200_plan_t* p) { return p ? p->m : 0; }                  \
  extern "C" int sdft_b200_##SFX##_window(const sdft_b200_plan_t* p) { return p ? p->window : 0; }              \
  extern "C" double sdft_b200_##SFX##_latency(const sdft_b200_plan_t* p) { return p ? p->latency : 0; }         \
  extern "C" void sdft_b200_##SFX##_sdft_n(sdft_b200_plan_t* p, size_t n, const TD* x, FDX* d)                  \
  {                                                                                                             \
    if (typed<TD, FD>(p, "sdft_sdft_n: plan type mismatch")) do_sdft<TD, FD>(p, n, x, (cx<FD>*)d);              \
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
  cudaSetDevice(p->device);
  cudaError_t e = cudaStreamSynchronize(p->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->copy_stream);
  unsigned ctl[2] = { 0, 0 };
  if (e == cudaSuccess) e = cudaMemcpy(ctl, p->control, sizeof(ctl), cudaMemcpyDeviceToHost);
  if (e != cudaSuccess) plan_fail(p, (int)e, "sdft_b200_synchronize", __FILE__, __LINE__);
  else if (ctl[1]) plan_fail(p, SDFT_B200_ERR_CHAIN, "chained scan: a carry wait timed out", __FILE__, __LINE__);
  return p->status;
}

extern "C" int sdft_b200_set_stream(sdft_b200_plan_t* p, void* cuda_stream)
{
  if (!p) return SDFT_B200_ERR_ARG;
  cudaStream_t next = (cuda_stream == (void*)-1) ? p->own_stream : (cudaStream_t)cuda_stream;
  if (next == p->stream) return p->status;
  cudaSetDevice(p->device);
  cudaStreamSynchronize(p->stream);   // work queued on the old stream must not race with the new one
  p->stream = next;
  return p->status;
}

extern "C" int sdft_b200_set_chunk(sdft_b200_plan_t* p, size_t chunk)
{
  if (!p) return SDFT_B200_ERR_ARG;
  p->forced_chunk = chunk;
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
  cudaSetDevice(p->device);
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
  cudaSetDevice(p->device);
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
extern "C" int sdft_b200_device(const sdft_b200_plan_t* p) { return p ? p->device : -1; }
extern "C" unsigned long long sdft_b200_launch_count(const sdft_b200_plan_t* p) { return p ? p->launches : 0; }

extern "C" int sdft_b200_get_twiddles(sdft_b200_plan_t* p, void* analysis, void* synthesis)
{
  if (!p) return SDFT_B200_ERR_ARG;
  cudaSetDevice(p->device);
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
  cudaSetDevice(p->device);
  const size_t cbytes = (p->fd == kF32) ? sizeof(cx<float>) : sizeof(cx<double>);
  const size_t tbytes = (p->td == kF32) ? sizeof(float) : sizeof(double);
  cudaError_t e = cudaStreamSynchronize(p->stream);
  if (cursor) *cursor = p->cursor;
  if (e == cudaSuccess && history)
    e = cudaMemcpy(history, (char*)p->history[p->hist_sel] + channel * 2 * p->m * tbytes, 2 * p->m * tbytes,
                   cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && accumulators)
  {
    e = cudaMemcpy(accumulators, (char*)p->acc_state[p->acc_sel] + (channel * p->cells + 2) * cbytes, p->m * cbytes,
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
      phase_at_kernel<float><<<blocks, threads, 0, p->stream>>>((const cx<float>*)p->tw_ext, (const cx<float>*)p->f0,
                                                                (cx<float>*)p->phase_scratch, (unsigned)p->cells, (unsigned)p->cursor);
    else
      phase_at_kernel<double><<<blocks, threads, 0, p->stream>>>((const cx<double>*)p->tw_ext, (const cx<double>*)p->f0,
                                                                 (cx<double>*)p->phase_scratch, (unsigned)p->cells, (unsigned)p->cursor);
    p->launches++;
    e = cudaStreamSynchronize(p->stream);
    if (e == cudaSuccess)
      e = cudaMemcpy(phase, (char*)p->phase_scratch + 2 * cbytes, p->m * cbytes, cudaMemcpyDeviceToHost);
  }
  if (e != cudaSuccess) plan_fail(p, (int)e, "sdft_b200_get_state", __FILE__, __LINE__);
  return p->status;
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

extern "C" const char* sdft_b200_version(void)
{
  return "sdft_b200 0.1 (sm_100a; analysis: chunked two-pass scan; synthesis: warp reduction)";
}
