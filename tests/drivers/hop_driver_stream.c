/*
 * C driver in the shape of the reference's test/test.c:49-93 with the buffers on the GPU and the plan in
 * STREAMING mode: plain C, no CUDA toolkit -- device memory comes from sdft_b200_device_alloc.  One plan, the signal
 * processed hop by hop through sdft_sdft_n (consecutive calls overlap on the device), then the whole matrix through
 * sdft_isdft_n; the DFT row of the first sample of every hop and the resynthesized samples go to stdout.
 * Input: raw float32 samples on stdin.  Output: nhops*dftsize complex128 rows, then nhops*hopsize float32 samples.
 * usage: hop_driver_stream dftsize hopsize window latency depth
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <sdft/sdft.h>

int main(int argc, char* argv[])
{
  if (argc < 6) return 1;
  const size_t dftsize = (size_t)atoi(argv[1]);
  const size_t hopsize = (size_t)atoi(argv[2]);
  const sdft_window_t window = (sdft_window_t)atoi(argv[3]);
  const double latency = atof(argv[4]);
  const unsigned depth = (unsigned)atoi(argv[5]);

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
  sdft_b200_plan_t* plan = (sdft_b200_plan_t*)sdft;

  /* everything the hop loop touches lives on the device; the samples are complete there before the first call
   * is issued (the streaming promise, include/sdft_b200.h) */
  float* d_input = (float*)sdft_b200_device_alloc(size * sizeof(float));
  float* d_output = (float*)sdft_b200_device_alloc(size * sizeof(float));
  sdft_fdx_t* d_rows = (sdft_fdx_t*)sdft_b200_device_alloc(size * dftsize * sizeof(sdft_fdx_t));
  if (!d_input || !d_output || !d_rows) return 3;
  if (sdft_b200_copy(plan, d_input, input, size * sizeof(float))) return 4;
  if (sdft_b200_set_streaming(plan, depth)) return 5;

  for (size_t i = 0; i < size; i += hopsize)
    sdft_sdft_n(sdft, hopsize, d_input + i, d_rows + i * dftsize);
  sdft_isdft_n(sdft, size, d_rows, d_output);          /* queued behind the calls: sees all of their rows */
  if (sdft_b200_synchronize(plan)) { fprintf(stderr, "%s\n", sdft_b200_last_error_string(plan)); return 6; }

  float* output = (float*)malloc(size * sizeof(float));
  sdft_fdx_t* dfts = (sdft_fdx_t*)malloc(nhops * dftsize * sizeof(sdft_fdx_t));
  if (sdft_b200_copy(plan, output, d_output, size * sizeof(float))) return 7;
  for (size_t j = 0; j < nhops; j++)
    if (sdft_b200_copy(plan, dfts + j * dftsize, d_rows + j * hopsize * dftsize, dftsize * sizeof(sdft_fdx_t))) return 8;

  fwrite(dfts, sizeof(sdft_fdx_t), nhops * dftsize, stdout);
  fwrite(output, sizeof(float), size, stdout);

  free(dfts); free(output); free(input);
  sdft_b200_device_free(d_rows); sdft_b200_device_free(d_output); sdft_b200_device_free(d_input);
  sdft_free(sdft);
  return 0;
}
