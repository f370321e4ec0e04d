// Dumps the twiddle tables of sdft_b200/csrc/sdft_tables.hpp (analysis then synthesis, interleaved re/im) to stdout:
//   tables_dump <f32|f64> <dftsize> <latency>
// tests/test_tables.py compares them bit for bit with the oracle's tables (c/src/sdft/sdft.h:439-446).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include "sdft_tables.hpp"

template <typename F> int dump(size_t m, double latency)
{
  std::vector<sdftb200::table_entry<F>> tw, tws;
  sdftb200::make_tables<F>(m, latency, tw, tws);
  fwrite(tw.data(), sizeof(tw[0]), m, stdout);
  fwrite(tws.data(), sizeof(tws[0]), m, stdout);
  return 0;
}

int main(int argc, char** argv)
{
  if (argc != 4) return 2;
  const size_t m = (size_t)atol(argv[2]);
  const double latency = atof(argv[3]);
  return strcmp(argv[1], "f32") == 0 ? dump<float>(m, latency) : dump<double>(m, latency);
}
