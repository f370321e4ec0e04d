/*
 * C++ driver in the shape of the reference's test/test.cpp: sdft::SDFT<float, double>, hop by hop.
 * Same stdin/stdout protocol as hop_driver.c.
 */
#include <complex>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include <sdft/sdft.h>

int main(int argc, char* argv[])
{
  if (argc < 5) return 1;
  const size_t dftsize = (size_t)atoi(argv[1]);
  const size_t hopsize = (size_t)atoi(argv[2]);
  const sdft::Window window = static_cast<sdft::Window>(atoi(argv[3]));
  const double latency = atof(argv[4]);

  std::vector<float> input;
  float chunk[4096];
  for (;;)
  {
    const size_t got = fread(chunk, sizeof(float), 4096, stdin);
    if (got == 0) break;
    input.insert(input.end(), chunk, chunk + got);
  }
  const size_t size = (input.size() / hopsize) * hopsize;
  const size_t nhops = size / hopsize;

  sdft::SDFT<float, double> sdft(dftsize, window, latency);
  if (sdft.size() != dftsize || sdft.window() != window || sdft.latency() != latency) return 3;

  std::vector<float> output(size);
  std::vector<std::complex<double>> buffer(hopsize * dftsize);
  std::vector<std::complex<double>> dfts(nhops * dftsize);

  for (size_t i = 0, j = 0; i < size; i += hopsize, j++)
  {
    sdft.sdft(hopsize, input.data() + i, buffer.data());
    sdft.isdft(hopsize, buffer.data(), output.data() + i);
    std::copy(buffer.begin(), buffer.begin() + dftsize, dfts.begin() + j * dftsize);
  }

  fwrite(dfts.data(), sizeof(std::complex<double>), dfts.size(), stdout);
  fwrite(output.data(), sizeof(float), output.size(), stdout);
  return 0;
}
