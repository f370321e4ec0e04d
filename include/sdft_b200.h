/*
 * sdft_b200.h -- C-ABI of libsdft_b200.so: the B200-native (sm_100a CUDA) replacement for the
 * analysis/synthesis hot path of jurihock/sdft.
 *
 * The library exports one symbol set per (time-domain, frequency-domain) type pair because the
 * reference selects those types at compile time (c/src/sdft/sdft.h:101-125) and reuses the same
 * function names for every pair.  Suffix <td><fd> is one of f32f32, f32f64 (the reference default),
 * f64f32, f64f64; long double has no device equivalent and is rejected by the shim header.
 * include/c/sdft/sdft.h maps the reference's public names onto these symbols, so existing C callers
 * only swap the include path and link -lsdft_b200.  include/cpp/sdft/sdft.h does the same for the
 * C++ template sdft::SDFT<T, F>.
 *
 * Pointer arguments (`samples`, `dfts`) may be HOST or DEVICE pointers; the library asks the CUDA
 * runtime which.  Host pointers: the call copies in/out and returns when the result is in place.
 * Device pointers: nothing is copied, work is queued on the plan's stream and the call returns at once
 * (stream-ordered; see sdft_b200_synchronize / sdft_b200_set_stream).
 *
 * Aliasing: *_roundtrip* may run in place (in == out); partially overlapping DEVICE ranges, and any overlap
 * of the DEVICE in/out of *_convolve_n, are rejected (error 10002 on the plan).  A call that returns while
 * its kernels are queued (device destination) has always finished reading the caller's HOST inputs.
 * Calls leave the calling thread's current CUDA device unchanged.
 *
 * There is no CPU fallback: if no CUDA device or no sm_100a image is usable, *_alloc* returns NULL and
 * sdft_b200_last_error_string(NULL) says why.
 *
 * Plain C, no CUDA or torch types in any signature.
 */
#ifndef SDFT_B200_H
#define SDFT_B200_H

#include <stddef.h>

#if defined(__cplusplus)
extern "C" {
#endif

#if defined(_WIN32)
#define SDFT_B200_API
#else
#define SDFT_B200_API __attribute__((visibility("default")))
#endif

/* interleaved {re, im}; layout-identical to the reference's sdft_fdx_t in all of its three spellings
 * (C99 complex, _Dcomplex, struct{r,i}: c/src/sdft/sdft.h:84-99), std::complex<F> and numpy complex */
typedef struct sdft_b200_cf32 { float r, i; } sdft_b200_cf32_t;
typedef struct sdft_b200_cf64 { double r, i; } sdft_b200_cf64_t;

/* same numbering as enum sdft_window (c/src/sdft/sdft.h:127-133) */
enum sdft_b200_window
{
  sdft_b200_window_boxcar = 0,
  sdft_b200_window_hann = 1,
  sdft_b200_window_hamming = 2,
  sdft_b200_window_blackman = 3
};

/* opaque device-resident plan (replaces struct sdft_plan, c/src/sdft/sdft.h:175-182) */
typedef struct sdft_b200_plan sdft_b200_plan_t;

/*
 * Drop-in entry points, one set per type pair.  Each replaces the reference function named in the
 * comment; argument meaning and ownership are unchanged (c/src/sdft/sdft.h line in brackets).
 */
#define SDFT_B200_DECLARE(SFX, TD, FDX)                                                                        \
  /* sdft_alloc [457]: hann, latency 1 */                                                                      \
  SDFT_B200_API sdft_b200_plan_t* sdft_b200_##SFX##_alloc(size_t dftsize);                                     \
  /* sdft_alloc_custom [413] */                                                                                \
  SDFT_B200_API sdft_b200_plan_t* sdft_b200_##SFX##_alloc_custom(size_t dftsize, int window, double latency);  \
  /* extension: `channels` independent plans advanced by one launch (SURVEY 8e channel sharding) */           \
  SDFT_B200_API sdft_b200_plan_t* sdft_b200_##SFX##_alloc_batch(size_t dftsize, int window, double latency,    \
                                                                size_t channels);                              \
  /* sdft_free [466]: NULL is a no-op */                                                                       \
  SDFT_B200_API void sdft_b200_##SFX##_free(sdft_b200_plan_t* plan);                                           \
  /* sdft_reset [517] */                                                                                       \
  SDFT_B200_API void sdft_b200_##SFX##_reset(sdft_b200_plan_t* plan);                                          \
  /* sdft_size [535], sdft_window [543], sdft_latency [551]: NULL -> 0 / boxcar / 0 */                         \
  SDFT_B200_API size_t sdft_b200_##SFX##_size(const sdft_b200_plan_t* plan);                                   \
  SDFT_B200_API int sdft_b200_##SFX##_window(const sdft_b200_plan_t* plan);                                    \
  SDFT_B200_API double sdft_b200_##SFX##_latency(const sdft_b200_plan_t* plan);                                \
  /* sdft_sdft [562] */                                                                                        \
  SDFT_B200_API void sdft_b200_##SFX##_sdft(sdft_b200_plan_t* plan, TD sample, FDX* dft);                      \
  /* sdft_sdft_n [607]: dfts is (nsamples, dftsize) row-major */                                               \
  SDFT_B200_API void sdft_b200_##SFX##_sdft_n(sdft_b200_plan_t* plan, size_t nsamples, const TD* samples,      \
                                              FDX* dfts);                                                      \
  /* sdft_sdft_nd [622]: dfts is nsamples row pointers */                                                      \
  SDFT_B200_API void sdft_b200_##SFX##_sdft_nd(sdft_b200_plan_t* plan, size_t nsamples, const TD* samples,     \
                                               FDX** dfts);                                                    \
  /* sdft_isdft [635] */                                                                                       \
  SDFT_B200_API TD sdft_b200_##SFX##_isdft(sdft_b200_plan_t* plan, const FDX* dft);                            \
  /* sdft_isdft_n [666] */                                                                                     \
  SDFT_B200_API void sdft_b200_##SFX##_isdft_n(sdft_b200_plan_t* plan, size_t nsamples, const FDX* dfts,       \
                                               TD* samples);                                                   \
  /* sdft_isdft_nd [681] */                                                                                    \
  SDFT_B200_API void sdft_b200_##SFX##_isdft_nd(sdft_b200_plan_t* plan, size_t nsamples, const FDX** dfts,     \
                                                TD* samples);                                                  \
  /* extension: analysis state update without writing rows (used to prime a time shard with its         \
   * 2m-sample halo) */                                                                             \
  SDFT_B200_API void sdft_b200_##SFX##_advance(sdft_b200_plan_t* plan, size_t nsamples, const TD* samples);    \
  /* extension: the hop loop of the reference's drivers (test/test.c:69-83) issued from inside the library:    \
   * exactly `for h < nhops: sdft_sdft_n(plan, hopsize, samples + h*hopsize, dfts + h*hop_stride)` with       \
   * hop_stride counted in complex values; saves the caller's per-call overhead when hops are short */         \
  SDFT_B200_API void sdft_b200_##SFX##_sdft_hops(sdft_b200_plan_t* plan, size_t nhops, size_t hopsize,         \
                                                 const TD* samples, FDX* dfts, size_t hop_stride);             \
  /* extension, batch plans: samples is (channels, nsamples), dfts is (channels, nsamples, dftsize) */         \
  SDFT_B200_API void sdft_b200_##SFX##_sdft_batch(sdft_b200_plan_t* plan, size_t nsamples, const TD* samples,  \
                                                  FDX* dfts);                                                  \
  SDFT_B200_API void sdft_b200_##SFX##_isdft_batch(sdft_b200_plan_t* plan, size_t nsamples, const FDX* dfts,   \
                                                   TD* samples);                                               \
  /* extension: fused analysis -> synthesis round trip (the test/test.c:79-80 pattern) that never       \
   * materialises the (n, m) matrix; out[t] equals isdft(sdft(in[t])) */                                       \
  SDFT_B200_API void sdft_b200_##SFX##_roundtrip_n(sdft_b200_plan_t* plan, size_t nsamples, const TD* in,      \
                                                   TD* out);                                                   \
  /* extension: SDFT.convolve of the reference's Python class (python/src/sdft/sdft.py:146-203): windows   \
   * `nsamples` UN-windowed rows in the frequency domain, out = window(in) / dftsize, mirror cells as in   \
   * sdft.h:589-595; stateless; in/out both (nsamples, dftsize), host or device */                          \
  SDFT_B200_API void sdft_b200_##SFX##_convolve_n(sdft_b200_plan_t* plan, size_t nsamples, const FDX* in,      \
                                                  FDX* out);                                                   \
  /* extension: the same with a spectral gain between analysis and synthesis: out[t] equals               \
   * isdft(gains .* sdft(in[t])), `gains` being dftsize complex factors (host or device memory) */            \
  SDFT_B200_API void sdft_b200_##SFX##_roundtrip_gain_n(sdft_b200_plan_t* plan, size_t nsamples, const TD* in, \
                                                        TD* out, const FDX* gains);

SDFT_B200_DECLARE(f32f32, float, sdft_b200_cf32_t)
SDFT_B200_DECLARE(f32f64, float, sdft_b200_cf64_t)
SDFT_B200_DECLARE(f64f32, double, sdft_b200_cf32_t)
SDFT_B200_DECLARE(f64f64, double, sdft_b200_cf64_t)

/*
 * Type-independent extensions (the plan remembers its type pair).
 */

/* 0 = ok; otherwise the (sticky) cudaError_t / library error code of the first failure on this plan.
 * plan == NULL reports the last error of a failed *_alloc* on the calling thread. */
SDFT_B200_API int sdft_b200_last_error(const sdft_b200_plan_t* plan);
SDFT_B200_API const char* sdft_b200_last_error_string(const sdft_b200_plan_t* plan);

/* Waits until everything queued on the plan's stream is done.  Returns sdft_b200_last_error. */
SDFT_B200_API int sdft_b200_synchronize(sdft_b200_plan_t* plan);

/* Makes the plan use a caller-owned stream (a cudaStream_t passed as void*; NULL = legacy default
 * stream).  The plan's own stream is kept for later sdft_b200_set_stream(plan, (void*)-1). */
SDFT_B200_API int sdft_b200_set_stream(sdft_b200_plan_t* plan, void* cuda_stream);

/* STREAMING MODE for endless sequences of short calls (the hop loop of test/test.c:69-83 with device buffers).
 * depth <= 1 (default): every call starts after the previous work of the stream has completed -- plain stream
 * order.  depth D > 1: up to D consecutive *_sdft_n / *_sdft_batch / *_advance calls with DEVICE samples and
 * DEVICE rows may be in flight at once: a call starts computing while its predecessors still stream their rows
 * out, and takes the history and the accumulators over through device-side counters instead of a kernel
 * boundary (a 4096-sample call at dftsize 512 otherwise spends most of its time on start-up and drain
 * latencies).  Results do not depend on timing; with the same chunk length they are bit-identical to the serial
 * mode (a streaming plan picks longer chunks by default, which changes the order in which carries are added:
 * ~1e-13 for double, ~1e-6 for float data).  What the caller promises while depth > 1:
 *   - the samples of a call are complete in device memory when the call is ISSUED (written by work that was
 *     synchronised with the host or finished before the previous library call was issued), not merely ordered
 *     before it on the stream;
 *   - the rows of a call are not overwritten by other work queued between two calls.
 * What still holds: calls COMPLETE in order, so anything queued behind them in the ordinary way (kernels,
 * copies, sdft_b200_synchronize, a stream or event wait) sees the rows and the state of all earlier calls.
 * Costs (depth + 1) copies of the per-channel state (history and accumulators) and `depth` sets of scan scratch.
 * Not for use from several host threads at once (as everything on one plan). */
SDFT_B200_API int sdft_b200_set_streaming(sdft_b200_plan_t* plan, unsigned depth);

/* Scan chunk length in samples (multiple of 32, <= 1024); 0 = choose per call from n and m. */
SDFT_B200_API int sdft_b200_set_chunk(sdft_b200_plan_t* plan, size_t chunk);

/* Region of interest (the reference plans carry one, c/src/sdft/sdft.h:137-143, but have no setter: always all
 * bins): from now on the rows written by *_sdft* and read by *_isdft* hold only bins [first, first + count),
 * i.e. they are (nsamples, count) matrices; count = 0 restores the whole spectrum.  Bins outside the region
 * still take part in the state update, they just cost no row bandwidth; *_isdft* of such rows sums the bins of
 * the region only.  The fused round trip and convolve always work on the whole spectrum. */
SDFT_B200_API int sdft_b200_set_roi(sdft_b200_plan_t* plan, size_t first, size_t count);

SDFT_B200_API size_t sdft_b200_channels(const sdft_b200_plan_t* plan);
SDFT_B200_API int sdft_b200_device(const sdft_b200_plan_t* plan);
/* device bytes of the plan's tables (twiddles, synthesis twiddles, phase source): O(dftsize) for a double
 * frequency domain, bounded by a fixed budget for float up to dftsize ~ 90 000 */
SDFT_B200_API size_t sdft_b200_table_bytes(const sdft_b200_plan_t* plan);
/* number of kernels this plan has launched so far (bench.py reports it as gpu_launches) */
SDFT_B200_API unsigned long long sdft_b200_launch_count(const sdft_b200_plan_t* plan);
/* number of analysis calls that ran as a wide body plus a narrow tail of bins in one launch (long float calls whose
 * last warp group would be mostly empty) -- introspection for the tests */
SDFT_B200_API unsigned long long sdft_b200_split_count(const sdft_b200_plan_t* plan);

/* CUDA-event timing of the dominant kernels on the plan's stream, for roofline reporting.
 * sdft_b200_kernel_ms returns the summed duration (ms) of the launches of kernel class `which`
 * (0 = analysis emit kernel, 1 = synthesis kernel) since the previous query, and their count. */
SDFT_B200_API int sdft_b200_set_profiling(sdft_b200_plan_t* plan, int on);
SDFT_B200_API double sdft_b200_kernel_ms(sdft_b200_plan_t* plan, int which, unsigned long long* launches);

/* Tuning aid: a library built with -DSDFT_B200_TRACE records eight %globaltimer stamps per CTA of the last
 * analysis launch (ticket, deltas loaded, chunk total, aggregate published, carry known, replay start,
 * rows done, unused); copies up to max_items x 8 stamps to `stamps` and returns the number of CTAs.
 * A regular build records nothing and returns 0. */
SDFT_B200_API size_t sdft_b200_debug_trace(sdft_b200_plan_t* plan, unsigned long long* stamps, size_t max_items);

/* Introspection used by the parity tests: copies plan tables/state to HOST buffers.
 * twiddles: analysis and synthesis tables, dftsize complex values each (c/src/sdft/sdft.h:444-445).
 * state (channel): cursor, history (2*dftsize samples, oldest first), accumulators and current
 * modulation phase (dftsize complex values each) (c/src/sdft/sdft.h:153-159). */
SDFT_B200_API int sdft_b200_get_twiddles(sdft_b200_plan_t* plan, void* analysis, void* synthesis);
SDFT_B200_API int sdft_b200_get_state(sdft_b200_plan_t* plan, size_t channel, size_t* cursor, void* history,
                                      void* accumulators, void* phase);

/* Import of plan state, the counterpart of sdft_b200_get_state: cursor (0 .. 2*dftsize-1), history
 * (2*dftsize samples, oldest first) and accumulators (dftsize complex values) from HOST buffers; a NULL
 * pointer leaves that part unchanged.  This is what exact time sharding uses: a shard
 * starts from the 2m samples before it and from the sum of the preceding shards' accumulator increments. */
SDFT_B200_API int sdft_b200_set_state(sdft_b200_plan_t* plan, size_t channel, size_t cursor, const void* history,
                                      const void* accumulators);

/* Roofline denominators of THIS board at THIS moment (bench.py): a pure streaming store with the row kernel's own
 * store instruction and cache policy (kind 0), a pure streaming read with the synthesis kernel's load (kind 1) and
 * a copy (kind 2, the buffer split in halves) over `bytes` of caller-provided device memory.  Returns GB/s of the
 * best of `reps` launches, the mean over all of them in *sustained (may be NULL); 0 on failure.  Nothing of the
 * transform is computed here.  sdft_b200_measure_dfma: DFMA instructions per second (per thread-lane) of a pure
 * FP64 FMA loop on every SM, the ceiling the fused round trip is quoted against. */
SDFT_B200_API double sdft_b200_measure_hbm(int kind, void* device_buffer, size_t bytes, int reps, double* sustained);
SDFT_B200_API double sdft_b200_measure_dfma(int reps);

/* Shard planners, pure integer arithmetic, the same as sdft_b200/shard.py:
 * sdft_b200_time_shard: rank `rank` of `world` analyses samples [*begin, *end) of an `nsamples`-sample signal after
 * priming a fresh plan with samples [*halo_begin, *begin) through *_advance (nothing to prime for rank 0); every
 * boundary is a multiple of 2*dftsize, where the reference's modulation phase restarts (c/src/sdft/sdft.h:566-576),
 * so every shard starts at cursor 0.  Trailing ranks may come out empty (*begin == *end) for short signals.
 * sdft_b200_channel_shard: contiguous channel blocks whose sizes differ by at most one; channels are independent
 * plans (sdft.h:175-180), nothing is exchanged.  Return 0, or 10002 for world == 0 / rank >= world. */
SDFT_B200_API int sdft_b200_time_shard(size_t nsamples, size_t world, size_t dftsize, size_t rank, size_t* begin,
                                       size_t* end, size_t* halo_begin);
SDFT_B200_API int sdft_b200_channel_shard(size_t channels, size_t world, size_t rank, size_t* begin, size_t* end);

/* Page-locked host memory so that host-pointer calls can DMA straight into the caller's buffer. */
SDFT_B200_API void* sdft_b200_host_alloc(size_t bytes);
SDFT_B200_API void sdft_b200_host_free(void* ptr);

/* Device memory for callers without the CUDA toolkit (plain C / C++ like the reference's drivers): a hop buffer
 * from sdft_b200_device_alloc instead of malloc keeps the rows on the GPU between sdft_sdft_n and sdft_isdft_n
 * (INTEGRATION.md).  sdft_b200_copy moves bytes between host and device memory in either direction, synchronously,
 * after the plan's queued work (plan may be NULL: no wait); returns 0 on success. */
SDFT_B200_API void* sdft_b200_device_alloc(size_t bytes);
SDFT_B200_API void sdft_b200_device_free(void* ptr);
SDFT_B200_API int sdft_b200_copy(sdft_b200_plan_t* plan, void* dst, const void* src, size_t bytes);

/* Library build information: "sdft_b200 <version> sm_100a ..." */
SDFT_B200_API const char* sdft_b200_version(void);

#if defined(__cplusplus)
}
#endif

#endif /* SDFT_B200_H */
