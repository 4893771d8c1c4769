// ipb200 — correctly rounded binary64 elementary functions for host and device.
//
// Why: the _d / _8 kinds of the reference evaluate their map projections with the platform's
// binary64 libm (src/ip_lambert_conf_grid_mod.F90:318-350, ip_polar_stereo_grid_mod.F90:344-443,
// ip_rot_equid_cylind_grid_mod.F90:361-436, movect.F90:37-64 ...).  Several of those formulas are
// ill-conditioned (movect divides by 1-crd**2 with crd up to 0.9999999, the rotated grids take acos
// of a value near 1), so a 1-ulp difference between two libms moves an output value by up to ~1e-9
// relative — a thousand times the 1e-12 parity budget.  libms differ (glibc 2.39 misrounds ~0.1 %
// of sin/cos/asin/atan2 arguments, CUDA's are 1-2 ulp), so the only platform-independent contract
// is *correct rounding*.  Every function here evaluates in double-double arithmetic (~2^-100
// relative) and rounds once, so host decode, device kernels and the test oracle (which gets the
// same contract independently, from libquadmath) agree bit for bit.
//
// Cost is irrelevant: these run only in the plan build (once per grid pair), never in the km-field loop.
// Out-of-domain / non-finite / huge arguments fall back to the platform function.
#pragma once
#include <cmath>

#if defined(__CUDACC__)
#define IPB_HD __host__ __device__ __forceinline__
#define IPB_HD_FN __host__ __device__ __noinline__ inline  /* entry points: one out-of-line copy per kernel image */
#else
#define IPB_HD inline
#define IPB_HD_FN inline
#endif

namespace ipb {
namespace cr {

struct dd { double h, l; };

// ---- error-free transformations (never contracted: intrinsics on device, -ffp-contract=off on host)
IPB_HD double add_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  return a + b;
#endif
}
IPB_HD double mul_(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  return a * b;
#endif
}
IPB_HD double fma_(double a, double b, double c) {
#ifdef __CUDA_ARCH__
  return __fma_rn(a, b, c);
#else
  return std::fma(a, b, c);
#endif
}
IPB_HD dd two_sum(double a, double b) {
  const double s = add_(a, b), bb = add_(s, -a);
  return {s, add_(add_(a, -add_(s, -bb)), add_(b, -bb))};
}
IPB_HD dd fast_two_sum(double a, double b) {  // |a| >= |b| or a == 0
  const double s = add_(a, b);
  return {s, add_(b, -add_(s, -a))};
}
IPB_HD dd two_prod(double a, double b) {
  const double p = mul_(a, b);
  return {p, fma_(a, b, -p)};
}

IPB_HD dd neg(dd a) { return {-a.h, -a.l}; }
IPB_HD dd add(dd a, dd b) {
  dd s = two_sum(a.h, b.h);
  const dd t = two_sum(a.l, b.l);
  s.l = add_(s.l, t.h);
  s = fast_two_sum(s.h, s.l);
  s.l = add_(s.l, t.l);
  return fast_two_sum(s.h, s.l);
}
IPB_HD dd add(dd a, double b) {
  dd s = two_sum(a.h, b);
  s.l = add_(s.l, a.l);
  return fast_two_sum(s.h, s.l);
}
IPB_HD dd sub(dd a, dd b) { return add(a, neg(b)); }
IPB_HD dd mul(dd a, dd b) {
  dd p = two_prod(a.h, b.h);
  p.l = add_(p.l, add_(mul_(a.h, b.l), mul_(a.l, b.h)));
  return fast_two_sum(p.h, p.l);
}
IPB_HD dd mul(dd a, double b) {
  dd p = two_prod(a.h, b);
  p.l = fma_(a.l, b, p.l);
  return fast_two_sum(p.h, p.l);
}
IPB_HD dd div(dd a, dd b) {
  const double q1 = a.h / b.h;
  dd r = sub(a, mul(b, q1));
  const double q2 = r.h / b.h;
  r = sub(r, mul(b, q2));
  const double q3 = r.h / b.h;
  return add(fast_two_sum(q1, q2), q3);
}
IPB_HD dd sqrt_dd(dd a) {
  if (!(a.h > 0)) return {a.h == 0 ? 0.0 : std::sqrt(a.h), 0.0};
  const double x = std::sqrt(a.h);
  const dd x2 = two_prod(x, x);
  const dd r = sub(a, x2);
  return fast_two_sum(x, r.h / (2 * x));
}
IPB_HD double rnd(dd a) { return add_(a.h, a.l); }

// ---- constants -----------------------------------------------------------------------------------
#define IPB_CR_PIO2_1 0x1.921fb54442d18p+0
#define IPB_CR_PIO2_2 0x1.1a62633145c07p-54
#define IPB_CR_PIO2_3 (-0x1.f1976b7ed8fbcp-110)
#define IPB_CR_LN2_1 0x1.62e42fefa39efp-1
#define IPB_CR_LN2_2 0x1.abc9e3b39803fp-56
#define IPB_CR_LN2_3 0x1.7b57a079a1934p-111

// 1/n! for n = 2..30 as double-double
IPB_HD dd inv_fact(int n) {
  switch (n) {
    case 2: return {0x1.0000000000000p-1, 0x0.0p+0};
    case 3: return {0x1.5555555555555p-3, 0x1.5555555555555p-57};
    case 4: return {0x1.5555555555555p-5, 0x1.5555555555555p-59};
    case 5: return {0x1.1111111111111p-7, 0x1.1111111111111p-63};
    case 6: return {0x1.6c16c16c16c17p-10, -0x1.f49f49f49f49fp-65};
    case 7: return {0x1.a01a01a01a01ap-13, 0x1.a01a01a01a01ap-73};
    case 8: return {0x1.a01a01a01a01ap-16, 0x1.a01a01a01a01ap-76};
    case 9: return {0x1.71de3a556c734p-19, -0x1.c154f8ddc6c00p-73};
    case 10: return {0x1.27e4fb7789f5cp-22, 0x1.cbbc05b4fa99ap-76};
    case 11: return {0x1.ae64567f544e4p-26, -0x1.c062e06d1f209p-80};
    case 12: return {0x1.1eed8eff8d898p-29, -0x1.2aec959e14c06p-83};
    case 13: return {0x1.6124613a86d09p-33, 0x1.f28e0cc748ebep-87};
    case 14: return {0x1.93974a8c07c9dp-37, 0x1.05d6f8a2efd1fp-92};
    case 15: return {0x1.ae7f3e733b81fp-41, 0x1.1d8656b0ee8cbp-97};
    case 16: return {0x1.ae7f3e733b81fp-45, 0x1.1d8656b0ee8cbp-101};
    case 17: return {0x1.952c77030ad4ap-49, 0x1.ac981465ddc6cp-103};
    case 18: return {0x1.6827863b97d97p-53, 0x1.eec01221a8b0bp-107};
    case 19: return {0x1.2f49b46814157p-57, 0x1.2650f61dbdcb4p-112};
    case 20: return {0x1.e542ba4020225p-62, 0x1.ea72b4afe3c2fp-120};
    case 21: return {0x1.71b8ef6dcf572p-66, -0x1.d043ae40c4647p-120};
    case 22: return {0x1.0ce396db7f853p-70, -0x1.aebcdbd20331cp-124};
    case 23: return {0x1.761b41316381ap-75, -0x1.3423c7d91404fp-130};
    case 24: return {0x1.f2cf01972f578p-80, -0x1.9ada5fcc1ab14p-135};
    case 25: return {0x1.3f3ccdd165fa9p-84, -0x1.58ddadf344487p-139};
    case 26: return {0x1.88e85fc6a4e5ap-89, -0x1.71c37ebd16540p-143};
    case 27: return {0x1.d1ab1c2dccea3p-94, 0x1.054d0c78aea14p-149};
    case 28: return {0x1.0a18a2635085dp-98, 0x1.b9e2e28e1aa54p-153};
    case 29: return {0x1.259f98b4358adp-103, 0x1.eaf8c39dd9bc5p-157};
    default: return {0x1.3932c5047d60ep-108, 0x1.832b7b530a627p-162};
  }
}

// ---- sin / cos -------------------------------------------------------------------------------------
// Taylor kernels on |r| <= pi/4 (+ a rounding error), r in double-double.
IPB_HD_FN void sincos_kernel(dd r, dd& s, dd& c) {
  const dd r2 = mul(r, r);
  // sin r = r (1 - r2/3! + r2^2/5! - ... + r2^14/29!)
  dd ps = inv_fact(29);
#pragma unroll
  for (int n = 27; n >= 3; n -= 2) ps = sub(inv_fact(n), mul(ps, r2));
  ps = sub(dd{1.0, 0.0}, mul(ps, r2));
  s = mul(ps, r);
  // cos r = 1 - r2/2! + r2^2/4! - ... + r2^15/30!
  dd pc = inv_fact(30);
#pragma unroll
  for (int n = 28; n >= 2; n -= 2) pc = sub(inv_fact(n), mul(pc, r2));
  c = sub(dd{1.0, 0.0}, mul(pc, r2));
}

// Argument reduction x = k*pi/2 + r for |x| < 2^25.  Returns false when x is outside that range.
IPB_HD bool reduce_pio2(double x, dd& r, int& q) {
  const double ax = x < 0 ? -x : x;
  if (!(ax < 33554432.0)) return false;
  if (ax <= 0.785398163397448279) { r = {x, 0.0}; q = 0; return true; }
  const double k = std::rint(mul_(x, 0x1.45f306dc9c883p-1));
  const dd p1 = two_prod(k, IPB_CR_PIO2_1), p2 = two_prod(k, IPB_CR_PIO2_2), p3 = two_prod(k, IPB_CR_PIO2_3);
  const double d = add_(x, -p1.h);  // exact (Sterbenz)
  dd acc = two_sum(d, -p1.l);
  acc = sub(acc, p2);
  acc = sub(acc, p3);
  r = acc;
  q = (int)((long long)k & 3);
  return true;
}

IPB_HD_FN bool sincos_dd(double x, dd& s, dd& c) {
  dd r; int q;
  if (!reduce_pio2(x, r, q)) return false;
  dd ks, kc;
  sincos_kernel(r, ks, kc);
  switch (q) {
    case 0: s = ks; c = kc; break;
    case 1: s = kc; c = neg(ks); break;
    case 2: s = neg(ks); c = neg(kc); break;
    default: s = neg(kc); c = ks; break;
  }
  return true;
}

IPB_HD_FN double sin(double x) {
  if (x == 0) return x;
  dd s, c;
  if (!sincos_dd(x, s, c)) return std::sin(x);
  return rnd(s);
}
IPB_HD_FN double cos(double x) {
  dd s, c;
  if (!sincos_dd(x, s, c)) return std::cos(x);
  return rnd(c);
}
IPB_HD_FN double tan(double x) {
  if (x == 0) return x;
  dd s, c;
  if (!sincos_dd(x, s, c)) return std::tan(x);
  return rnd(div(s, c));
}

// ---- exp / log / pow ---------------------------------------------------------------------------------
// exp of a double-double, |x| < 700.  Result as double-double (not scaled when k is returned).
IPB_HD_FN dd exp_dd(dd x) {
  const double k = std::rint(mul_(x.h, 0x1.71547652b82fep+0));
  dd r = sub(x, two_prod(k, IPB_CR_LN2_1));
  r = sub(r, two_prod(k, IPB_CR_LN2_2));
  r = sub(r, two_prod(k, IPB_CR_LN2_3));
  r = {mul_(r.h, 0.03125), mul_(r.l, 0.03125)};  // /32: |r| <= 0.0109
  // expm1(r) = r + r^2/2! + ... + r^14/14!
  dd p = inv_fact(14);
#pragma unroll
  for (int n = 13; n >= 2; n--) p = add(inv_fact(n), mul(p, r));
  p = add(dd{1.0, 0.0}, mul(p, r));
  dd s = mul(p, r);  // expm1(r)
#pragma unroll
  for (int i = 0; i < 5; i++) s = add(dd{mul_(s.h, 2.0), mul_(s.l, 2.0)}, mul(s, s));  // expm1(2r) = 2s + s^2
  s = add(s, 1.0);
  const int ki = (int)k;
  return {std::ldexp(s.h, ki), std::ldexp(s.l, ki)};
}
IPB_HD_FN double exp(double x) {
  if (!(x > -700.0 && x < 700.0)) return std::exp(x);
  if (x == 0) return 1.0;
  return rnd(exp_dd(dd{x, 0.0}));
}
// log of a positive, finite, normal double as double-double: one Newton step on exp from the
// platform's estimate, y1 = y0 + log(x e^-y0), |x e^-y0 - 1| ~ 2^-52.
IPB_HD_FN dd log_dd(double x) {
  const double y0 = std::log(x);
  const dd e = exp_dd(dd{-y0, 0.0});
  const dd t = add(mul(e, x), -1.0);
  const double t2 = mul_(t.h, t.h);
  dd y = add(t, mul_(-0.5, t2));
  return add(y, y0);
}
IPB_HD bool log_ok(double x) { return x >= 2.2250738585072014e-308 && x <= 1.7976931348623157e+308; }
IPB_HD_FN double log(double x) {
  if (!log_ok(x)) return std::log(x);
  if (x == 1.0) return 0.0;
  return rnd(log_dd(x));
}
IPB_HD_FN double pow(double a, double b) {
  if (!log_ok(a) || !(b == b) || b > 1e300 || b < -1e300) return std::pow(a, b);
  if (b == 0 || a == 1.0) return 1.0;
  const dd p = mul(log_dd(a), b);
  if (!(p.h > -700.0 && p.h < 700.0)) return std::pow(a, b);
  return rnd(exp_dd(p));
}

// ---- inverse trigonometric -----------------------------------------------------------------------------
// atan of a double-double |q| <= ~1e150: platform estimate refined by one exact-tangent step,
// atan q = y0 + atan((q cos y0 - sin y0) / (cos y0 + q sin y0)), the correction being ~2^-52.
IPB_HD_FN dd atan_dd(dd q) {
  const double y0 = std::atan(q.h);
  dd s, c;
  sincos_dd(y0, s, c);
  const dd num = sub(mul(q, c), s), den = add(c, mul(q, s));
  const dd d = div(num, den);
  return add(d, y0);
}
IPB_HD dd pi_dd() { return {0x1.921fb54442d18p+1, 0x1.1a62633145c07p-53}; }
IPB_HD dd pio2_dd() { return {IPB_CR_PIO2_1, IPB_CR_PIO2_2}; }
IPB_HD bool mid_range(double a) { const double x = a < 0 ? -a : a; return x > 1e-140 && x < 1e140; }

IPB_HD_FN double atan(double x) {
  if (x == 0 || !mid_range(x)) return std::atan(x);
  const double ax = x < 0 ? -x : x;
  dd a;
  if (ax <= 1.0) a = atan_dd(dd{x, 0.0});
  else {
    a = atan_dd(div(dd{1.0, 0.0}, dd{x, 0.0}));
    a = sub(x > 0 ? pio2_dd() : neg(pio2_dd()), a);
  }
  return rnd(a);
}
// atan2 with double-double operands; both non-zero, moderate magnitude.
IPB_HD_FN dd atan2_dd(dd y, dd x) {
  const double ay = y.h < 0 ? -y.h : y.h, ax = x.h < 0 ? -x.h : x.h;
  if (ax >= ay) {
    dd a = atan_dd(div(y, x));
    if (x.h < 0) a = add(a, y.h >= 0 ? pi_dd() : neg(pi_dd()));
    return a;
  }
  const dd a = atan_dd(div(x, y));
  return sub(y.h > 0 ? pio2_dd() : neg(pio2_dd()), a);
}
IPB_HD_FN double atan2(double y, double x) {
  if (y == 0 || x == 0 || !mid_range(y) || !mid_range(x)) return std::atan2(y, x);
  return rnd(atan2_dd(dd{y, 0.0}, dd{x, 0.0}));
}
IPB_HD_FN double asin(double x) {
  const double ax = x < 0 ? -x : x;
  if (!(ax <= 1.0) || ax < 1e-140) return std::asin(x);
  if (ax == 1.0) return x > 0 ? IPB_CR_PIO2_1 : -IPB_CR_PIO2_1;
  const dd t = sqrt_dd(mul(two_sum(1.0, -x), two_sum(1.0, x)));  // sqrt(1-x^2) > 0
  return rnd(atan2_dd(dd{x, 0.0}, t));
}
IPB_HD_FN double acos(double x) {
  const double ax = x < 0 ? -x : x;
  if (!(ax <= 1.0)) return std::acos(x);
  if (x == 1.0) return 0.0;
  if (x == -1.0) return 0x1.921fb54442d18p+1;
  if (ax < 1e-140) return IPB_CR_PIO2_1;
  const dd t = sqrt_dd(mul(two_sum(1.0, -x), two_sum(1.0, x)));
  return rnd(atan2_dd(t, dd{x, 0.0}));
}

}  // namespace cr
}  // namespace ipb
