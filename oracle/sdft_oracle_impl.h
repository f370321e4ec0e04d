/*
 * oracle/sdft_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the modulated sliding DFT of jurihock/sdft, instantiated once per
 * (time-domain, frequency-domain) type pair by sdft_oracle.c.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference leg may load the library built from this file.
 *
 * Parity: PINNED.  tests/test_oracle_vs_ref.py checks this restatement bit-for-bit against the
 * unmodified reference header compiled under oracle/_ref (see oracle/Makefile), and
 * tests/test_oracle_golden.py checks it against committed vectors generated from that build and from
 * the reference's Python class (tests/golden/make_golden.py).
 *
 * The restatement is organised around the *scan* view of the algorithm rather
 * than the reference's per-sample plan mutation:
 *   - history is a "last 2m samples" delay line with its own head index, exported oldest-first
 *     (the reference indexes the ring by cursor, c/src/sdft/sdft.h:564; the two are equivalent);
 *   - bins live in an *extended* index space e = k + 2, e in [0, m+4), whose four outer cells are the
 *     mirror cells of c/src/sdft/sdft.h:589-595; re/im are kept in separate arrays;
 *   - the modulation phase ("fiddle") is P[cursor][k], restarted at 1 every 2m samples
 *     (c/src/sdft/sdft.h:566-576).
 * Every floating-point expression keeps the operand order and the rounding points of the reference
 * (cited per function) so that results are bit-identical for all four type pairs.
 *
 * Required macros: OR_TD (time-domain type), OR_FD (frequency-domain type), OR_SUFFIX (symbol suffix),
 * OR_COS / OR_SIN / OR_ACOS (libm functions of the FD precision, c/src/sdft/sdft.h:193-213,333-348).
 */

#define OR_CAT2(a, b) a##b
#define OR_CAT(a, b) OR_CAT2(a, b)
#define OR_FN(name) OR_CAT(OR_CAT(oracle_, OR_SUFFIX), OR_CAT(_, name))
#define OR_PLAN OR_CAT(oracle_plan_, OR_SUFFIX)

typedef struct OR_PLAN
{
  size_t m;            /* number of bins (dftsize) */
  int window;          /* 0 boxcar, 1 hann, 2 hamming, 3 blackman (c/src/sdft/sdft.h:127-133) */
  double latency;      /* synthesis latency factor, kept in double (sdft.h:166) */
  OR_FD scale;         /* analysis weight 1/(2m) (sdft.h:422) */
  size_t cursor;       /* position inside the 2m period, 0..2m-1 (sdft.h:153) */
  size_t head;         /* index of the oldest sample inside history[] */
  OR_TD* history;      /* last 2m samples, circular, oldest at head */
  OR_FD* tw_re;        /* analysis twiddle, per bin (sdft.h:444) */
  OR_FD* tw_im;
  OR_FD* syn_re;       /* synthesis twiddle, per bin (sdft.h:445) */
  OR_FD* syn_im;
  OR_FD* acc_re;       /* running modulated sums, per bin (sdft.h:157) */
  OR_FD* acc_im;
  OR_FD* ph_re;        /* current modulation phase P[cursor], per bin (sdft.h:159) */
  OR_FD* ph_im;
  OR_FD* ext_re;       /* demodulated spectrum in extended bin space, m+4 cells (sdft.h:158) */
  OR_FD* ext_im;
} OR_PLAN;

/* Table generation: c/src/sdft/sdft.h:439-446.  Type notes: omega, omega*i and
 * omega*i*m are FD-typed products; the trailing "* latency" is a double product that is narrowed to
 * FD only when handed to cos/sin. */
static void OR_FN(tables)(OR_PLAN* p)
{
  const size_t m = p->m;
  const OR_FD omega = (OR_FD)(-2) * OR_ACOS((OR_FD)(-1)) / (OR_FD)(m * 2);
  const OR_FD warg = (OR_FD)((omega * (OR_FD)m) * p->latency);
  const OR_FD wsyn = (OR_FD)(+2) / ((OR_FD)(1) - OR_COS(warg));
  for (size_t k = 0; k < m; ++k)
  {
    const OR_FD a = omega * (OR_FD)k;
    const OR_FD unit = (OR_FD)(1);
    p->tw_re[k] = unit * OR_COS(a);
    p->tw_im[k] = unit * OR_SIN(a);
    const OR_FD s = (OR_FD)(((omega * (OR_FD)k) * (OR_FD)m) * p->latency);
    p->syn_re[k] = wsyn * OR_COS(s);
    p->syn_im[k] = wsyn * OR_SIN(s);
  }
}

/* c/src/sdft/sdft.h:517-529 */
void OR_FN(reset)(OR_PLAN* p)
{
  const size_t m = p->m;
  p->cursor = 0;
  p->head = 0;
  memset(p->history, 0, 2 * m * sizeof(OR_TD));
  memset(p->acc_re, 0, m * sizeof(OR_FD));
  memset(p->acc_im, 0, m * sizeof(OR_FD));
  memset(p->ext_re, 0, (m + 4) * sizeof(OR_FD));
  memset(p->ext_im, 0, (m + 4) * sizeof(OR_FD));
  for (size_t k = 0; k < m; ++k)
  {
    p->ph_re[k] = (OR_FD)1;
    p->ph_im[k] = (OR_FD)0;
  }
}

/* c/src/sdft/sdft.h:413-450 */
OR_PLAN* OR_FN(alloc)(size_t m, int window, double latency)
{
  OR_PLAN* p = (OR_PLAN*)calloc(1, sizeof(OR_PLAN));
  p->m = m;
  p->window = window;
  p->latency = latency;
  p->scale = (OR_FD)(1) / (OR_FD)(m * 2);
  p->history = (OR_TD*)calloc(2 * m, sizeof(OR_TD));
  OR_FD** arrays[] = { &p->tw_re, &p->tw_im, &p->syn_re, &p->syn_im,
                       &p->acc_re, &p->acc_im, &p->ph_re, &p->ph_im };
  for (size_t i = 0; i < sizeof(arrays) / sizeof(arrays[0]); ++i)
  {
    *arrays[i] = (OR_FD*)calloc(m, sizeof(OR_FD));
  }
  p->ext_re = (OR_FD*)calloc(m + 4, sizeof(OR_FD));
  p->ext_im = (OR_FD*)calloc(m + 4, sizeof(OR_FD));
  OR_FN(tables)(p);
  OR_FN(reset)(p);
  return p;
}

/* c/src/sdft/sdft.h:466-511 */
void OR_FN(free)(OR_PLAN* p)
{
  if (!p) return;
  free(p->history);
  free(p->tw_re); free(p->tw_im); free(p->syn_re); free(p->syn_im);
  free(p->acc_re); free(p->acc_im); free(p->ph_re); free(p->ph_im);
  free(p->ext_re); free(p->ext_im);
  free(p);
}

/* Window as a 1/3/5-tap combination of neighbouring extended cells: c/src/sdft/sdft.h:350-402.
 * c is the centre cell index in extended space; returns one component. */
static inline OR_FD OR_FN(tap)(const OR_FD* x, size_t c, int window, OR_FD scale)
{
  switch (window)
  {
    case 1: /* hann, sdft.h:366-374: ((mid+mid) - (l+r)) * (w*0.25) */
    {
      const OR_FD a = x[c] + x[c];
      const OR_FD b = x[c - 1] + x[c + 1];
      return (a - b) * (scale * (OR_FD)(0.25));
    }
    case 2: /* hamming, sdft.h:375-383 */
    {
      const OR_FD a = x[c] * (OR_FD)(0.54);
      const OR_FD b = (x[c - 1] + x[c + 1]) * (OR_FD)(0.23);
      return (a - b) * scale;
    }
    case 3: /* blackman, sdft.h:384-393 */
    {
      const OR_FD a = x[c] * (OR_FD)(0.42);
      const OR_FD b = (x[c - 1] + x[c + 1]) * (OR_FD)(0.25);
      const OR_FD d = (x[c - 2] + x[c + 2]) * (OR_FD)(0.04);
      return ((a - b) + d) * scale;
    }
    default: /* boxcar, sdft.h:394-399 */
      return x[c] * scale;
  }
}

/* One analysis step: c/src/sdft/sdft.h:562-598.  out = m interleaved (re, im) pairs. */
static void OR_FN(step)(OR_PLAN* p, OR_TD sample, OR_FD* out)
{
  const size_t m = p->m;
  const size_t period = 2 * m;

  /* oldest sample leaves, newest enters; the difference is formed in TD precision (sdft.h:564) */
  const OR_TD oldest = p->history[p->head];
  p->history[p->head] = sample;
  p->head = (p->head + 1 == period) ? 0 : p->head + 1;
  const OR_TD tdelta = sample - oldest;
  const OR_FD delta = (OR_FD)tdelta;

  const int wrap = (p->cursor >= period - 1);
  p->cursor = wrap ? 0 : p->cursor + 1;

  OR_FD* xr = p->ext_re + 2;
  OR_FD* xi = p->ext_im + 2;

  for (size_t k = 0; k < m; ++k)
  {
    /* acc += P * delta  (sdft.h:572 / :583) */
    const OR_FD tr = p->ph_re[k] * delta;
    const OR_FD ti = p->ph_im[k] * delta;
    p->acc_re[k] = p->acc_re[k] + tr;
    p->acc_im[k] = p->acc_im[k] + ti;

    if (wrap)
    {
      /* phase restarts, spectrum is the accumulator itself (sdft.h:573-574) */
      p->ph_re[k] = (OR_FD)1;
      p->ph_im[k] = (OR_FD)0;
      xr[k] = p->acc_re[k];
      xi[k] = p->acc_im[k];
    }
    else
    {
      /* P *= tw (sdft.h:584 with the product formula of :298-300) */
      const OR_FD pr = p->ph_re[k], pi = p->ph_im[k];
      const OR_FD wr = p->tw_re[k], wi = p->tw_im[k];
      const OR_FD nr = pr * wr - pi * wi;
      const OR_FD ni = pr * wi + pi * wr;
      p->ph_re[k] = nr;
      p->ph_im[k] = ni;
      /* spectrum = acc * conj(P)  (sdft.h:585) */
      const OR_FD cr = nr, ci = -ni;
      xr[k] = p->acc_re[k] * cr - p->acc_im[k] * ci;
      xi[k] = p->acc_re[k] * ci + p->acc_im[k] * cr;
    }
  }

  /* mirror cells (sdft.h:589-595): below bin 0 about bin 0, above bin m-1 about bin m-1 (sic).
   * Same interleaved write order as the reference loop, which matters for m < 3. */
  for (size_t i = 1; i <= 2; ++i)
  {
    p->ext_re[2 - i] = p->ext_re[2 + i];
    p->ext_im[2 - i] = -p->ext_im[2 + i];
    p->ext_re[(m + 1) + i] = p->ext_re[(m + 1) - i];
    p->ext_im[(m + 1) + i] = -p->ext_im[(m + 1) - i];
  }

  for (size_t k = 0; k < m; ++k)
  {
    out[2 * k + 0] = OR_FN(tap)(p->ext_re, k + 2, p->window, p->scale);
    out[2 * k + 1] = OR_FN(tap)(p->ext_im, k + 2, p->window, p->scale);
  }
}

/* c/src/sdft/sdft.h:607-613; dfts is (n, m) row-major interleaved complex */
void OR_FN(sdft_n)(OR_PLAN* p, size_t n, const OR_TD* samples, OR_FD* dfts)
{
  for (size_t t = 0; t < n; ++t)
  {
    OR_FN(step)(p, samples[t], dfts + t * 2 * p->m);
  }
}

/* c/src/sdft/sdft.h:635-657 */
static OR_TD OR_FN(synth)(const OR_PLAN* p, const OR_FD* dft)
{
  const size_t m = p->m;
  OR_FD y = (OR_FD)0;
  if (p->latency == 1)
  {
    for (size_t k = 0; k < m; ++k)
    {
      y += dft[2 * k] * (OR_FD)((k % 2) ? -1 : +1);
    }
  }
  else
  {
    for (size_t k = 0; k < m; ++k)
    {
      y += dft[2 * k] * p->syn_re[k] - dft[2 * k + 1] * p->syn_im[k];
    }
  }
  y *= (OR_FD)2;
  return (OR_TD)y;
}

/* c/src/sdft/sdft.h:666-672 */
void OR_FN(isdft_n)(OR_PLAN* p, size_t n, const OR_FD* dfts, OR_TD* samples)
{
  for (size_t t = 0; t < n; ++t)
  {
    samples[t] = OR_FN(synth)(p, dfts + t * 2 * p->m);
  }
}

/* ---- helpers for the full-size parity tests (tests/test_gpu_configs.py) ----
 * The CPU cannot afford to produce every row of a 2^20 x 4096 matrix, so the tests walk the state
 * through the whole signal and produce rows only at sampled positions.  None of these change the
 * arithmetic: `advance_n` performs exactly the state updates of one analysis step (sdft.h:564-587)
 * and skips the demodulation / mirror / window stages, which only feed the output row. */
void OR_FN(advance_n)(OR_PLAN* p, size_t n, const OR_TD* samples)
{
  const size_t m = p->m;
  const size_t period = 2 * m;
  for (size_t t = 0; t < n; ++t)
  {
    const OR_TD sample = samples[t];
    const OR_TD oldest = p->history[p->head];
    p->history[p->head] = sample;
    p->head = (p->head + 1 == period) ? 0 : p->head + 1;
    const OR_TD tdelta = sample - oldest;                 /* sdft.h:564 */
    const OR_FD delta = (OR_FD)tdelta;
    const int wrap = (p->cursor >= period - 1);
    p->cursor = wrap ? 0 : p->cursor + 1;
    for (size_t k = 0; k < m; ++k)
    {
      const OR_FD tr = p->ph_re[k] * delta;               /* sdft.h:572 / :583 */
      const OR_FD ti = p->ph_im[k] * delta;
      p->acc_re[k] = p->acc_re[k] + tr;
      p->acc_im[k] = p->acc_im[k] + ti;
      if (wrap)
      {
        p->ph_re[k] = (OR_FD)1;                           /* sdft.h:573 */
        p->ph_im[k] = (OR_FD)0;
      }
      else
      {
        const OR_FD pr = p->ph_re[k], pi = p->ph_im[k];   /* sdft.h:584 */
        const OR_FD wr = p->tw_re[k], wi = p->tw_im[k];
        p->ph_re[k] = pr * wr - pi * wi;
        p->ph_im[k] = pr * wi + pi * wr;
      }
    }
  }
}

/* deep copy of a plan, optionally with another window (the window only enters the output stage,
 * sdft.h:597, so plans that differ in nothing but the window share their state evolution) */
OR_PLAN* OR_FN(clone)(const OR_PLAN* src, int window)
{
  OR_PLAN* p = OR_FN(alloc)(src->m, window < 0 ? src->window : window, src->latency);
  const size_t m = src->m;
  p->cursor = src->cursor;
  p->head = src->head;
  memcpy(p->history, src->history, 2 * m * sizeof(OR_TD));
  memcpy(p->acc_re, src->acc_re, m * sizeof(OR_FD));
  memcpy(p->acc_im, src->acc_im, m * sizeof(OR_FD));
  memcpy(p->ph_re, src->ph_re, m * sizeof(OR_FD));
  memcpy(p->ph_im, src->ph_im, m * sizeof(OR_FD));
  memcpy(p->ext_re, src->ext_re, (m + 4) * sizeof(OR_FD));
  memcpy(p->ext_im, src->ext_im, (m + 4) * sizeof(OR_FD));
  return p;
}

/* analysis immediately followed by synthesis, sample by sample (the test/test.c:79-80 pattern with a
 * hop of one): y[t] = isdft(sdft(x[t])).  `row` is caller scratch of 2m OR_FD values. */
void OR_FN(roundtrip_n)(OR_PLAN* p, size_t n, const OR_TD* samples, OR_TD* out, OR_FD* row)
{
  for (size_t t = 0; t < n; ++t)
  {
    OR_FN(step)(p, samples[t], row);
    out[t] = OR_FN(synth)(p, row);
  }
}

/* ---- introspection used by the table/state parity tests ---- */

size_t OR_FN(size)(const OR_PLAN* p) { return p ? p->m : 0; }
size_t OR_FN(cursor)(const OR_PLAN* p) { return p->cursor; }

/* interleaved (re, im), m entries each */
void OR_FN(get_twiddles)(const OR_PLAN* p, OR_FD* analysis, OR_FD* synthesis)
{
  for (size_t k = 0; k < p->m; ++k)
  {
    analysis[2 * k] = p->tw_re[k];  analysis[2 * k + 1] = p->tw_im[k];
    synthesis[2 * k] = p->syn_re[k]; synthesis[2 * k + 1] = p->syn_im[k];
  }
}

/* history: 2m samples oldest first; acc, phase: interleaved, m entries each */
void OR_FN(get_state)(const OR_PLAN* p, OR_TD* history, OR_FD* acc, OR_FD* phase)
{
  for (size_t i = 0; i < 2 * p->m; ++i)
  {
    history[i] = p->history[(p->head + i) % (2 * p->m)];
  }
  for (size_t k = 0; k < p->m; ++k)
  {
    acc[2 * k] = p->acc_re[k];   acc[2 * k + 1] = p->acc_im[k];
    phase[2 * k] = p->ph_re[k];  phase[2 * k + 1] = p->ph_im[k];
  }
}

#undef OR_CAT2
#undef OR_CAT
#undef OR_FN
#undef OR_PLAN
