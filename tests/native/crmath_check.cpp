// Host check of csrc/ipb_crmath.h (product) and oracle/ip_oracle_libm.hpp (oracle) against libquadmath (binary128 evaluation rounded once to binary64).
// Prints one line per function: name, samples, mismatches.  Exit code = number of functions with mismatches.
// Build: g++ -O2 -std=c++17 -ffp-contract=off crmath_check.cpp -lquadmath
#include <quadmath.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>

#include "../../nceplibs-ip_b200/csrc/ipb_crmath.h"
#include "../../oracle/ip_oracle_libm.hpp"

using namespace ipb;
static bool same(double a, double b) { return std::memcmp(&a, &b, 8) == 0 || (a != a && b != b); }

int main(int argc, char** argv) {
  const long N = argc > 1 ? atol(argv[1]) : 200000;
  std::mt19937_64 rng(20261019);
  auto U = [&](double a, double b) { return std::uniform_real_distribution<double>(a, b)(rng); };
  auto logU = [&](double e0, double e1) { return std::ldexp(U(1, 2), (int)U(e0, e1)) * (rng() & 1 ? 1 : -1); };
  int bad_fn = 0;
  struct Row { const char* name; long n, bad; double worst_x, worst_y; };
  auto report = [&](Row r) {
    printf("%-6s samples %ld mismatches %ld", r.name, r.n, r.bad);
    if (r.bad) printf("  e.g. x=%a y=%a", r.worst_x, r.worst_y);
    printf("\n");
    if (r.bad) bad_fn++;
  };
#define CHECK1(NAME, GEN, QF)                                                     \
  {                                                                               \
    Row r{#NAME, 0, 0, 0, 0};                                                     \
    for (long i = 0; i < N; i++) {                                                \
      const double x = GEN;                                                       \
      const double got = cr::NAME(x), ref = (double)QF((__float128)x);            \
      r.n++;                                                                      \
      if (!same(got, ref) || !same(ipo::lm::NAME(x), ref)) { r.bad++; r.worst_x = x; } \
    }                                                                             \
    report(r);                                                                    \
  }
  // arguments as the projections produce them: angles in radians up to a few turns, plus wide sweeps
  CHECK1(sin, (i & 1) ? U(-7, 7) : logU(-60, 24), sinq)
  CHECK1(cos, (i & 1) ? U(-7, 7) : logU(-60, 24), cosq)
  CHECK1(tan, (i & 1) ? U(-1.6, 1.6) : logU(-60, 24), tanq)
  CHECK1(asin, (i & 1) ? U(-1, 1) : (1 - std::ldexp(U(1, 2), (int)U(-52, -1))) * (rng() & 1 ? 1 : -1), asinq)
  CHECK1(acos, (i & 1) ? U(-1, 1) : (1 - std::ldexp(U(1, 2), (int)U(-52, -1))) * (rng() & 1 ? 1 : -1), acosq)
  CHECK1(atan, (i & 1) ? U(-4, 4) : logU(-60, 60), atanq)
  CHECK1(exp, (i & 1) ? U(-20, 20) : logU(-40, 9), expq)
  CHECK1(log, (i & 1) ? U(1e-3, 100) : std::fabs(logU(-300, 300)), logq)
  {
    Row r{"atan2", 0, 0, 0, 0};
    for (long i = 0; i < N; i++) {
      const double y = (i & 1) ? U(-3, 3) : logU(-40, 40), x = (i & 2) ? U(-3, 3) : logU(-40, 40);
      const double got = cr::atan2(y, x), ref = (double)atan2q((__float128)y, (__float128)x);
      r.n++;
      if (!same(got, ref) || !same(ipo::lm::atan2(y, x), ref)) { r.bad++; r.worst_x = x; r.worst_y = y; }
    }
    report(r);
  }
  {
    Row r{"pow", 0, 0, 0, 0};
    for (long i = 0; i < N; i++) {
      const double a = (i & 1) ? U(1e-3, 50) : std::fabs(logU(-30, 30)), b = (i & 2) ? U(-3, 3) : logU(-10, 4);
      const double got = cr::pow(a, b), ref = (double)powq((__float128)a, (__float128)b);
      r.n++;
      if (!same(got, ref) || !same(ipo::lm::pow(a, b), ref)) { r.bad++; r.worst_x = a; r.worst_y = b; }
    }
    report(r);
  }
  return bad_fn;
}
