// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the shipped product.
//
// The binary64 elementary functions the oracle's restatement of the reference calls.
//
// The reference calls the platform libm (gfortran -> glibc, ifort -> libimf), and several of its formulas
// amplify a 1-ulp libm difference to ~1e-9 relative in the output (movect.F90:44-52 divides by 1-crd**2;
// ip_rot_equid_cylind_grid_mod.F90:378,436 take acos near 1), so "the reference's bits" are defined only
// up to the libm.  Two flavours are therefore built from this one source (oracle/Makefile):
//
//   libip_oracle.so        default: every function CORRECTLY ROUNDED — the platform-independent
//                          contract the CUDA library is held to at the 1e-12 / bit level.  Implemented
//                          here independently of the product: x87 long double (64-bit significand, glibc
//                          ldbl-96 routines, < 1 ulp) with a Ziv test on the 11 guard bits, falling
//                          back to libquadmath's binary128 routine when the guard bits are within 2 ulp
//                          of a rounding boundary.
//   libip_oracle_glibc.so  -DORACLE_LIBM_GLIBC: plain glibc binary64 calls, i.e. what a gfortran build
//                          of the reference does on this host.  Used for the CPU timing baseline and to
//                          measure how far libm choice alone moves the reference's results
//                          (tests/test_libm_sensitivity.py).
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

#ifndef ORACLE_LIBM_GLIBC
#include <quadmath.h>
#endif

namespace ipo {
namespace lm {

#ifdef ORACLE_LIBM_GLIBC
inline double sin(double x) { return ::sin(x); }
inline double cos(double x) { return ::cos(x); }
inline double tan(double x) { return ::tan(x); }
inline double asin(double x) { return ::asin(x); }
inline double acos(double x) { return ::acos(x); }
inline double atan(double x) { return ::atan(x); }
inline double atan2(double y, double x) { return ::atan2(y, x); }
inline double log(double x) { return ::log(x); }
inline double exp(double x) { return ::exp(x); }
inline double pow(double a, double b) { return ::pow(a, b); }
inline const char* flavour() { return "glibc"; }
#else
// true when rounding the 64-bit-significand value y to binary64 cannot be affected by an error of
// up to 2 units in its last place: the 11 bits below the binary64 significand are not within 2 of
// the half-way pattern 0x400
inline bool safe_to_round(long double y) {
  if (!(y == y) || y == 0) return true;
  const long double ay = y < 0 ? -y : y;
  if (ay < 1e-290L || ay > 1e300L) return false;
  unsigned char b[16];
  std::memcpy(b, &y, sizeof(long double) < 16 ? sizeof(long double) : 16);
  uint64_t m;
  std::memcpy(&m, b, 8);  // x87 extended: explicit 64-bit significand in the low 8 bytes
  const int g = (int)(m & 0x7ff);
  return g < 0x400 - 2 || g > 0x400 + 2;
}
#define IPO_CR1(NAME)                                              \
  inline double NAME(double x) {                                   \
    const long double y = ::NAME##l((long double)x);               \
    if (safe_to_round(y)) return (double)y;                        \
    return (double)::NAME##q((__float128)x);                       \
  }
IPO_CR1(sin)
IPO_CR1(cos)
IPO_CR1(tan)
IPO_CR1(asin)
IPO_CR1(acos)
IPO_CR1(atan)
IPO_CR1(log)
IPO_CR1(exp)
#undef IPO_CR1
inline double atan2(double y, double x) {
  const long double r = ::atan2l((long double)y, (long double)x);
  if (safe_to_round(r)) return (double)r;
  return (double)::atan2q((__float128)y, (__float128)x);
}
inline double pow(double a, double b) {
  const long double r = ::powl((long double)a, (long double)b);
  if (safe_to_round(r)) return (double)r;
  return (double)::powq((__float128)a, (__float128)b);
}
inline const char* flavour() { return "correctly-rounded"; }
#endif

}  // namespace lm
}  // namespace ipo
