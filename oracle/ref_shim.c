/*
 * oracle/ref_shim.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin export layer around the UNMODIFIED reference header, which is compiled from where it lies
 * (-I/root/reference/c/src, see oracle/Makefile); nothing of it is copied into this repository.
 * One shared object per (TD, FD) pair because the header defines non-static functions whose ABI
 * depends on SDFT_TD_x / SDFT_FD_x (c/src/sdft/sdft.h:101-125).  -DSDFT_NO_COMPLEX_H is mandatory
 * under gcc.
 *
 * Exported: ref_alloc, ref_free, ref_reset, ref_size, ref_window, ref_latency, ref_sdft_n,
 * ref_isdft_n, ref_sdft_nd, ref_isdft_nd, ref_get_twiddles, ref_get_state, ref_td_size, ref_fd_size.
 */
#include <sdft/sdft.h>

#define REF_API __attribute__((visibility("default")))

REF_API void* ref_alloc(size_t m, int window, double latency)
{
  return sdft_alloc_custom(m, (sdft_window_t)window, latency);
}
REF_API void* ref_alloc_default(size_t m) { return sdft_alloc(m); }
REF_API void ref_free(void* p) { sdft_free((sdft_t*)p); }
REF_API void ref_reset(void* p) { sdft_reset((sdft_t*)p); }
REF_API size_t ref_size(const void* p) { return sdft_size((const sdft_t*)p); }
REF_API int ref_window(const void* p) { return (int)sdft_window((const sdft_t*)p); }
REF_API double ref_latency(const void* p) { return sdft_latency((const sdft_t*)p); }
REF_API size_t ref_td_size(void) { return sizeof(sdft_td_t); }
REF_API size_t ref_fd_size(void) { return sizeof(sdft_fd_t); }

REF_API void ref_sdft_n(void* p, size_t n, const sdft_td_t* x, sdft_fdx_t* dfts)
{
  sdft_sdft_n((sdft_t*)p, n, x, dfts);
}
REF_API void ref_isdft_n(void* p, size_t n, const sdft_fdx_t* dfts, sdft_td_t* y)
{
  sdft_isdft_n((sdft_t*)p, n, dfts, y);
}
REF_API void ref_sdft_nd(void* p, size_t n, const sdft_td_t* x, sdft_fdx_t** dfts)
{
  sdft_sdft_nd((sdft_t*)p, n, x, dfts);
}
REF_API void ref_isdft_nd(void* p, size_t n, const sdft_fdx_t** dfts, sdft_td_t* y)
{
  sdft_isdft_nd((sdft_t*)p, n, dfts, y);
}

/* interleaved (re, im), m entries each */
REF_API void ref_get_twiddles(const void* p, sdft_fd_t* analysis, sdft_fd_t* synthesis)
{
  const sdft_t* s = (const sdft_t*)p;
  for (size_t k = 0; k < s->dftsize; ++k)
  {
    analysis[2 * k] = s->analysis.twiddles[k].r;
    analysis[2 * k + 1] = s->analysis.twiddles[k].i;
    synthesis[2 * k] = s->synthesis.twiddles[k].r;
    synthesis[2 * k + 1] = s->synthesis.twiddles[k].i;
  }
}

/* history linearised oldest-first (the ring slot at cursor holds the oldest sample) */
REF_API size_t ref_get_state(const void* p, sdft_td_t* history, sdft_fd_t* acc, sdft_fd_t* phase)
{
  const sdft_t* s = (const sdft_t*)p;
  const size_t period = 2 * s->dftsize;
  for (size_t i = 0; i < period; ++i)
  {
    history[i] = s->analysis.input[(s->analysis.cursor + i) % period];
  }
  for (size_t k = 0; k < s->dftsize; ++k)
  {
    acc[2 * k] = s->analysis.accoutput[k].r;
    acc[2 * k + 1] = s->analysis.accoutput[k].i;
    phase[2 * k] = s->analysis.fiddles[k].r;
    phase[2 * k + 1] = s->analysis.fiddles[k].i;
  }
  return s->analysis.cursor;
}
