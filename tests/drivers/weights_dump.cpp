// Dumps the folded synthesis weights of sdft_b200/csrc/sdft_weights.hpp for tests/test_weights.py:
//   stdin : int64 m, int64 window, double prescale, double vr[m], double vi[m]
//   stdout: double ab[2m] (A, B per bin), then one byte: 1 when every B is zero
#include <cstdint>
#include <cstdio>
#include <vector>
#include "sdft_weights.hpp"

int main()
{
  int64_t head[2];
  double prescale;
  if (fread(head, sizeof(head), 1, stdin) != 1 || fread(&prescale, sizeof(prescale), 1, stdin) != 1) return 2;
  const size_t m = (size_t)head[0];
  std::vector<double> vr(m), vi(m);
  if (fread(vr.data(), sizeof(double), m, stdin) != m || fread(vi.data(), sizeof(double), m, stdin) != m) return 2;
  std::vector<double> ab;
  const bool unit = sdftb200::synth_weights<double>(m, (int)head[1], sdftb200::make_mirrors(m), prescale, vr.data(), vi.data(), ab);
  fwrite(ab.data(), sizeof(double), ab.size(), stdout);
  const unsigned char u = unit ? 1 : 0;
  fwrite(&u, 1, 1, stdout);
  return 0;
}
