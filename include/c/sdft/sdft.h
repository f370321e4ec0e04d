/*
 * sdft/sdft.h (C) -- drop-in replacement for the header of jurihock/sdft (c/src/sdft/sdft.h) that runs
 * the analysis/synthesis hot path on a B200 through libsdft_b200.so.
 *
 * Usage is unchanged: select the data types with the reference's macros before including,
 *
 *     #define SDFT_TD_FLOAT      // default      (or SDFT_TD_DOUBLE)
 *     #define SDFT_FD_DOUBLE     // default      (or SDFT_FD_FLOAT)
 *     #include <sdft/sdft.h>
 *
 * put this directory (include/c) on the include path instead of the reference's c/src, and link
 * -lsdft_b200.  Every public function of the reference exists with the same name, arguments and
 * ownership rules; each forwards to the type-suffixed symbol of include/sdft_b200.h that matches
 * the selected macros (the reference resolves the same choice at compile time, sdft.h:101-125).
 *
 * Differences a caller can observe:
 *   - sdft_t is opaque (the reference's struct fields are not part of its documented API);
 *   - `samples` / `dfts` may also be DEVICE pointers, in which case nothing is copied;
 *   - SDFT_TD_LONG_DOUBLE / SDFT_FD_LONG_DOUBLE are rejected: there is no 80-bit arithmetic on the GPU;
 *   - on a machine without a usable CUDA device sdft_alloc* returns NULL (no CPU fallback).
 */
#ifndef SDFT_B200_C_SHIM_H
#define SDFT_B200_C_SHIM_H

#include <stddef.h>

#include "../../sdft_b200.h"

#if defined(SDFT_TD_LONG_DOUBLE) || defined(SDFT_FD_LONG_DOUBLE)
#error "sdft_b200: long double time/frequency domain types are not supported on the GPU"
#endif

#if !defined(SDFT_NO_COMPLEX_H) && !defined(__cplusplus) && !defined(_MSC_VER)
#include <complex.h>
#define SDFT_B200_HAVE_C99_COMPLEX
#endif

#if defined(__cplusplus)
extern "C" {
#endif

typedef size_t sdft_size_t;
typedef float sdft_float_t;
typedef double sdft_double_t;

/* complex spellings of the reference (sdft.h:84-99); all are interleaved {re, im} */
#if defined(SDFT_B200_HAVE_C99_COMPLEX)
typedef float complex sdft_float_complex_t;
typedef double complex sdft_double_complex_t;
#else
struct sdft_float_complex { float r, i; };
struct sdft_double_complex { double r, i; };
typedef struct sdft_float_complex sdft_float_complex_t;
typedef struct sdft_double_complex sdft_double_complex_t;
#endif

#if defined(SDFT_TD_DOUBLE)
typedef sdft_double_t sdft_td_t;
#define SDFT_B200_TD f64
#else
#if !defined(SDFT_TD_FLOAT)
#define SDFT_TD_FLOAT
#endif
typedef sdft_float_t sdft_td_t;
#define SDFT_B200_TD f32
#endif

#if defined(SDFT_FD_FLOAT)
typedef sdft_float_t sdft_fd_t;
typedef sdft_float_complex_t sdft_fdx_t;
typedef sdft_b200_cf32_t sdft_b200_fdx_abi_t;
#define SDFT_B200_FD f32
#else
#if !defined(SDFT_FD_DOUBLE)
#define SDFT_FD_DOUBLE
#endif
typedef sdft_double_t sdft_fd_t;
typedef sdft_double_complex_t sdft_fdx_t;
typedef sdft_b200_cf64_t sdft_b200_fdx_abi_t;
#define SDFT_B200_FD f64
#endif

enum sdft_window
{
  sdft_window_boxcar,
  sdft_window_hann,
  sdft_window_hamming,
  sdft_window_blackman
};
typedef enum sdft_window sdft_window_t;

typedef sdft_b200_plan_t sdft_t;

#define SDFT_B200_PASTE3(a, b, c) a##b##c
#define SDFT_B200_PASTE(a, b, c) SDFT_B200_PASTE3(a, b, c)
#define SDFT_B200_FN(name) SDFT_B200_PASTE(SDFT_B200_PASTE(sdft_b200_, SDFT_B200_TD, SDFT_B200_FD), _, name)

static inline sdft_t* sdft_alloc(const sdft_size_t dftsize)
{
  return SDFT_B200_FN(alloc)(dftsize);
}

static inline sdft_t* sdft_alloc_custom(const sdft_size_t dftsize, const sdft_window_t window, const sdft_double_t latency)
{
  return SDFT_B200_FN(alloc_custom)(dftsize, (int)window, latency);
}

static inline void sdft_free(sdft_t* sdft) { SDFT_B200_FN(free)(sdft); }
static inline void sdft_reset(sdft_t* sdft) { SDFT_B200_FN(reset)(sdft); }
static inline sdft_size_t sdft_size(const sdft_t* sdft) { return SDFT_B200_FN(size)(sdft); }
static inline sdft_window_t sdft_window(const sdft_t* sdft) { return (sdft_window_t)SDFT_B200_FN(window)(sdft); }
static inline sdft_double_t sdft_latency(const sdft_t* sdft) { return SDFT_B200_FN(latency)(sdft); }

static inline void sdft_sdft(sdft_t* sdft, const sdft_td_t sample, sdft_fdx_t* const dft)
{
  SDFT_B200_FN(sdft)(sdft, sample, (sdft_b200_fdx_abi_t*)dft);
}

static inline void sdft_sdft_n(sdft_t* sdft, const sdft_size_t nsamples, const sdft_td_t* samples, sdft_fdx_t* const dfts)
{
  SDFT_B200_FN(sdft_n)(sdft, nsamples, samples, (sdft_b200_fdx_abi_t*)dfts);
}

static inline void sdft_sdft_nd(sdft_t* sdft, const sdft_size_t nsamples, const sdft_td_t* samples, sdft_fdx_t** const dfts)
{
  SDFT_B200_FN(sdft_nd)(sdft, nsamples, samples, (sdft_b200_fdx_abi_t**)dfts);
}

static inline sdft_td_t sdft_isdft(sdft_t* sdft, const sdft_fdx_t* dft)
{
  return SDFT_B200_FN(isdft)(sdft, (const sdft_b200_fdx_abi_t*)dft);
}

static inline void sdft_isdft_n(sdft_t* sdft, const sdft_size_t nsamples, const sdft_fdx_t* dfts, sdft_td_t* const samples)
{
  SDFT_B200_FN(isdft_n)(sdft, nsamples, (const sdft_b200_fdx_abi_t*)dfts, samples);
}

static inline void sdft_isdft_nd(sdft_t* sdft, const sdft_size_t nsamples, const sdft_fdx_t** dfts, sdft_td_t* const samples)
{
  SDFT_B200_FN(isdft_nd)(sdft, nsamples, (const sdft_b200_fdx_abi_t**)dfts, samples);
}

#if defined(__cplusplus)
}
#endif

#endif /* SDFT_B200_C_SHIM_H */
