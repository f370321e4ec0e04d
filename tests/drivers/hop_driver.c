/*
 * C driver in the shape of the reference's test/test.c:49-93: one plan, the signal processed hop by
 * hop through sdft_sdft_n + sdft_isdft_n, the DFT row of the first sample of every hop kept.
 * Input: raw float32 samples on stdin.  Output (stdout): nhops*dftsize complex128 rows, then
 * nhops*hopsize float32 resynthesized samples.  Types are the reference defaults (float / double).
 * usage: hop_driver dftsize hopsize window latency
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sdft/sdft.h>

int main(int argc, char* argv[])
{
  if (argc < 5) return 1;
  const size_t dftsize = (size_t)atoi(argv[1]);
  const size_t hopsize = (size_t)atoi(argv[2]);
  const sdft_window_t window = (sdft_window_t)atoi(argv[3]);
  const double latency = atof(argv[4]);

  size_t cap = 1 << 16, size = 0;
  float* input = (float*)malloc(cap * sizeof(float));
  for (;;)
  {
    if (size == cap) { cap *= 2; input = (float*)realloc(input, cap * sizeof(float)); }
    const size_t got = fread(input + size, sizeof(float), cap - size, stdin);
    if (got == 0) break;
    size += got;
  }
  size = (size / hopsize) * hopsize;
  const size_t nhops = size / hopsize;

  sdft_t* sdft = sdft_alloc_custom(dftsize, window, latency);
  if (!sdft) { fprintf(stderr, "no plan: %s\n", sdft_b200_last_error_string(NULL)); return 2; }
  if (sdft_size(sdft) != dftsize || sdft_window(sdft) != window || sdft_latency(sdft) != latency) return 3;

  float* output = (float*)malloc(size * sizeof(float));
  sdft_fdx_t* buffer = (sdft_fdx_t*)malloc(hopsize * dftsize * sizeof(sdft_fdx_t));
  sdft_fdx_t* dfts = (sdft_fdx_t*)malloc(nhops * dftsize * sizeof(sdft_fdx_t));

  for (size_t i = 0, j = 0; i < size; i += hopsize, j++)
  {
    sdft_sdft_n(sdft, hopsize, input + i, buffer);
    sdft_isdft_n(sdft, hopsize, buffer, output + i);
    memcpy(dfts + j * dftsize, buffer, dftsize * sizeof(sdft_fdx_t));
  }

  fwrite(dfts, sizeof(sdft_fdx_t), nhops * dftsize, stdout);
  fwrite(output, sizeof(float), size, stdout);

  free(dfts); free(buffer); free(output); free(input);
  sdft_free(sdft);
  return 0;
}
