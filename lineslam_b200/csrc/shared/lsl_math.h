// lsl_math.h — portable IEEE-754 double transcendental functions shared by the
// CPU oracle and the sm_100a kernels.
//
// Why this exists: LSD (external/lsd/lsd.cpp) takes
// hard decisions on the results of atan2/sin/cos/exp/log/log10/pow/sinh
// (gaussian_kernel lsd.cpp:466-489, isaligned :799-832, nfa :980-1065, region_grow
// :1638-1655). glibc libm and CUDA libdevice differ by 1-2 ulp, which is enough
// to flip a decision (measured: 725 vs 729 segments on chairs.pgm with a 1-ulp
// exp), so both sides evaluate THE SAME sequence of IEEE add/mul/div/sqrt
// operations. Host builds use -ffp-contract=off, device builds --fmad=false;
// neither compiler re-associates, so results are bit-identical by construction.
//
// Accuracy: every public function refines a ~1-ulp seed in double-double
// arithmetic (Dekker/Knuth error-free transforms, no FMA needed) and rounds
// once, i.e. the result is the correctly rounded value except when the exact
// value lies within ~1e-12 ulp of a rounding boundary. glibc 2.39's functions
// are themselves < 0.52 ulp, so the oracle built on these functions reproduces
// the unmodified upstream lsd.c + glibc segment lists (tests/test_oracle_ref.py).
#pragma once
#include <stdint.h>
#include <string.h>
#include <math.h>

#if defined(__CUDACC__)
#define LSL_HD __host__ __device__ __forceinline__
#define LSL_HDN static __host__ __device__ __noinline__
#else
#define LSL_HD static inline
#define LSL_HDN static inline
#endif

namespace lslm {

LSL_HD uint64_t d2u(double x) {
#if defined(__CUDA_ARCH__)
  return (uint64_t)__double_as_longlong(x);
#else
  uint64_t u; memcpy(&u, &x, 8); return u;
#endif
}
LSL_HD double u2d(uint64_t u) {
#if defined(__CUDA_ARCH__)
  return __longlong_as_double((long long)u);
#else
  double x; memcpy(&x, &u, 8); return x;
#endif
}

#define LSL_PI      3.14159265358979323846
#define LSL_PI_LO   1.2246467991473531772e-16
#define LSL_PIO2    1.57079632679489655800e+00
#define LSL_PIO2_LO 6.12323399573676603587e-17
#define LSL_LN10    2.30258509299404568402
#define LSL_NAN     (lslm::u2d(0x7ff8000000000000ull))

LSL_HD double pow2i(int k) { return u2d((uint64_t)(k + 1023) << 52); }
LSL_HD double scale2(double y, int k) {  // y * 2^k without libm
  if (k >= -1021 && k <= 1023) return y * pow2i(k);
  if (k > 1023) return y * pow2i(1023) * pow2i(k - 1023 > 1023 ? 1023 : k - 1023);
  int k2 = k + 1000; if (k2 < -1021) k2 = -1021;
  return y * pow2i(k2) * pow2i(-1000);
}

// ------------------------------------------------------- double-double ----
struct dd { double h, l; };
LSL_HD dd mkdd(double h, double l) { dd r; r.h = h; r.l = l; return r; }
LSL_HD dd two_sum(double a, double b) {
  double s = a + b, bb = s - a;
  return mkdd(s, (a - (s - bb)) + (b - bb));
}
LSL_HD dd fast_two_sum(double a, double b) {  // requires |a| >= |b| (or a == 0)
  double s = a + b;
  return mkdd(s, b - (s - a));
}
// p + e = a * b exactly (error-free product). Host: Dekker / Veltkamp splitting (no FMA is assumed: the reference
// build has none). Device: one explicit fused multiply-add — the error term of a product is representable, so both
// forms return the same (p, e) bit for bit for every product whose halves neither overflow nor underflow.
LSL_HD dd two_prod(double a, double b) {
  double p = a * b;
#if defined(__CUDA_ARCH__)
  return mkdd(p, __fma_rn(a, b, -p));
#endif
  double t = 134217729.0 * a, ah = t - (t - a), al = a - ah;
  t = 134217729.0 * b; double bh = t - (t - b), bl = b - bh;
  return mkdd(p, ((ah * bh - p) + ah * bl + al * bh) + al * bl);
}
LSL_HD dd dd_add(dd a, dd b) {
  dd s = two_sum(a.h, b.h), t = two_sum(a.l, b.l);
  s.l += t.h; s = fast_two_sum(s.h, s.l);
  s.l += t.l; return fast_two_sum(s.h, s.l);
}
LSL_HD dd dd_add_d(dd a, double b) {
  dd s = two_sum(a.h, b);
  s.l += a.l; return fast_two_sum(s.h, s.l);
}
LSL_HD dd dd_neg(dd a) { return mkdd(-a.h, -a.l); }
LSL_HD dd dd_mul(dd a, dd b) {
  dd p = two_prod(a.h, b.h);
  p.l += a.h * b.l + a.l * b.h;
  return fast_two_sum(p.h, p.l);
}
LSL_HD dd dd_mul_d(dd a, double b) {
  dd p = two_prod(a.h, b);
  p.l += a.l * b;
  return fast_two_sum(p.h, p.l);
}
LSL_HD dd dd_div(dd a, dd b) {
  double q1 = a.h / b.h;
  dd r = dd_add(a, dd_neg(dd_mul_d(b, q1)));
  double q2 = r.h / b.h;
  r = dd_add(r, dd_neg(dd_mul_d(b, q2)));
  double q3 = r.h / b.h;
  dd q = fast_two_sum(q1, q2);
  return dd_add_d(q, q3);
}

// ------------------------------------------------------------ ~1-ulp seeds ----
LSL_HD double seed_log(double x) {  // x normal, positive
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01,
               Lg3 = 2.857142874366239149e-01, Lg4 = 2.222219843214978396e-01,
               Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
  uint64_t u = d2u(x);
  int k = (int)(u >> 52) - 1023;
  uint64_t m = u & 0x000fffffffffffffull;
  if (m >= 0x6a09e667f3bcdull) { k += 1; x = u2d(m | 0x3fe0000000000000ull); }
  else x = u2d(m | 0x3ff0000000000000ull);
  double f = x - 1.0, s = f / (2.0 + f), dk = (double)k;
  double z = s * s, w = z * z;
  double t1 = w * (Lg2 + w * (Lg4 + w * Lg6));
  double t2 = z * (Lg1 + w * (Lg3 + w * (Lg5 + w * Lg7)));
  double R = t2 + t1, hfsq = 0.5 * f * f;
  return dk * ln2_hi - ((hfsq - (s * (hfsq + R) + dk * ln2_lo)) - f);
}
LSL_HD double seed_atan(double x) {
  const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01,
               aT2 = 1.42857142725034663711e-01, aT3 = -1.11111104054623557880e-01,
               aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
               aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02,
               aT8 = 4.97687799461593236017e-02, aT9 = -3.65315727442169155270e-02,
               aT10 = 1.62858201153657823623e-02;
  double ax = fabs(x), hi = 0.0, lo = 0.0; int id = -1;
  if (ax >= 7.378697629483821e19) { double r = LSL_PIO2 + LSL_PIO2_LO; return x > 0 ? r : -r; }
  if (ax < 0.4375) { if (ax < 1.862645149230957e-09) return x; }
  else if (ax < 1.1875) {
    if (ax < 0.6875) { id = 0; ax = (2.0 * ax - 1.0) / (2.0 + ax);
      hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17; }
    else { id = 1; ax = (ax - 1.0) / (ax + 1.0);
      hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17; }
  } else {
    if (ax < 2.4375) { id = 2; ax = (ax - 1.5) / (1.0 + 1.5 * ax);
      hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17; }
    else { id = 3; ax = -1.0 / ax;
      hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17; }
  }
  double z = ax * ax, w = z * z;
  double s1 = z * (aT0 + w * (aT2 + w * (aT4 + w * (aT6 + w * (aT8 + w * aT10)))));
  double s2 = w * (aT1 + w * (aT3 + w * (aT5 + w * (aT7 + w * aT9))));
  if (id < 0) { double r = ax - ax * (s1 + s2); return x < 0 ? -r : r; }
  double r = hi - ((ax * (s1 + s2) - lo) - ax);
  return x < 0 ? -r : r;
}
LSL_HD double seed_atan2(double y, double x) {  // finite, non-zero x and y
  int sy = (int)(d2u(y) >> 63), sx = (int)(d2u(x) >> 63);
  int m = sy | (sx << 1);
  double ax = fabs(x), ay = fabs(y), z;
  int k = (int)((d2u(ay) >> 52) & 0x7ff) - (int)((d2u(ax) >> 52) & 0x7ff);
  if (k > 60) z = LSL_PIO2 + 0.5 * LSL_PI_LO;
  else if (sx && k < -60) z = 0.0;
  else z = seed_atan(ay / ax);
  switch (m) {
    case 0: return z;
    case 1: return -z;
    case 2: return LSL_PI - (z - LSL_PI_LO);
    default: return (z - LSL_PI_LO) - LSL_PI;
  }
}

// ------------------------------------------------------------------ exp ----
// exp of a double-double argument, |x| < 745; result relative error ~1e-29.
LSL_HDN dd exp_dd(dd x, int* kout) {
  const double invln2 = 1.44269504088896338700e+00;
  const dd LN2 = mkdd(6.93147180559945286e-01, 2.31904681384629956e-17);
  double kf = floor(x.h * invln2 + 0.5);
  dd r = dd_add(x, dd_neg(dd_mul_d(LN2, kf)));
  r.h *= 0.001953125; r.l *= 0.001953125;  // / 512, exact
  // expm1(r) = r (1 + r/2! + r^2/3! + ... + r^8/9!)
  dd p = mkdd(2.75573192239858925e-06, -1.85839327404647208e-22);          // 1/9!
  p = dd_add(dd_mul(p, r), mkdd(2.48015873015873016e-05, 2.15119478667758816e-23));
  p = dd_add(dd_mul(p, r), mkdd(1.98412698412698413e-04, 1.72095582934207053e-22));
  p = dd_add(dd_mul(p, r), mkdd(1.38888888888888894e-03, -5.30054395437357706e-20));
  p = dd_add(dd_mul(p, r), mkdd(8.33333333333333322e-03, 1.15648231731787138e-19));
  p = dd_add(dd_mul(p, r), mkdd(4.16666666666666644e-02, 2.31296463463574266e-18));
  p = dd_add(dd_mul(p, r), mkdd(1.66666666666666657e-01, 9.25185853854297066e-18));
  p = dd_add(dd_mul(p, r), mkdd(0.5, 0.0));
  p = dd_add_d(dd_mul(p, r), 1.0);
  dd s = dd_mul(p, r);
  for (int i = 0; i < 9; ++i) {  // expm1(2a) = 2 expm1(a) + expm1(a)^2
    dd s2 = dd_mul(s, s);
    s = dd_add(mkdd(2.0 * s.h, 2.0 * s.l), s2);
  }
  *kout = (int)kf;
  return dd_add_d(s, 1.0);
}
LSL_HD double lsl_exp(double x) {
  if (x != x) return x;
  if (x > 7.09782712893383973096e+02) return HUGE_VAL;
  if (x < -7.45133219101941108420e+02) return 0.0;
  if (fabs(x) < 5.5e-17) return 1.0 + x;
  int k; dd e = exp_dd(mkdd(x, 0.0), &k);
  return scale2(e.h, k);
}

// ------------------------------------------------------------------ log ----
// log(x) as double-double: seed + one Newton step on exp.
LSL_HDN dd log_dd(double x) {  // x positive, normal
  // scale into [1,2) x 2^k by hand so exp_dd never sees a huge argument
  uint64_t u = d2u(x);
  int k = (int)(u >> 52) - 1023;
  uint64_t m = u & 0x000fffffffffffffull;
  double xm;
  if (m >= 0x6a09e667f3bcdull) { k += 1; xm = u2d(m | 0x3fe0000000000000ull); }
  else xm = u2d(m | 0x3ff0000000000000ull);
  double y0 = seed_log(xm);
  int ke; dd e = exp_dd(mkdd(-y0, 0.0), &ke);
  e.h = scale2(e.h, ke); e.l = scale2(e.l, ke);     // |ke| <= 1 here
  dd t = dd_add_d(dd_mul_d(e, xm), -1.0);            // xm*exp(-y0) - 1  (~1e-16)
  dd lm = dd_add(mkdd(y0, 0.0), t);                  // log(xm), t^2/2 negligible (1e-32)
  const dd LN2 = mkdd(6.93147180559945286e-01, 2.31904681384629956e-17);
  return dd_add(dd_mul_d(LN2, (double)k), lm);
}
LSL_HD int log_special(double* x, double* res, int* kadj) {
  *kadj = 0;
  if (*x != *x) { *res = *x; return 1; }
  if (*x < 0.0) { *res = LSL_NAN; return 1; }
  if (*x == 0.0) { *res = -HUGE_VAL; return 1; }
  if (*x == HUGE_VAL) { *res = *x; return 1; }
  if ((d2u(*x) >> 52) == 0) { *x *= 18014398509481984.0; *kadj = -54; }
  return 0;
}
LSL_HD double lsl_log(double x) {
  double r; int kadj;
  if (log_special(&x, &r, &kadj)) return r;
  if (x == 1.0) return 0.0;
  dd l = log_dd(x);
  if (kadj) l = dd_add(l, dd_mul_d(mkdd(6.93147180559945286e-01, 2.31904681384629956e-17), (double)kadj));
  return l.h;
}
LSL_HD double lsl_log10(double x) {
  double r; int kadj;
  if (log_special(&x, &r, &kadj)) return r;
  if (x == 1.0) return 0.0;
  dd l = log_dd(x);
  if (kadj) l = dd_add(l, dd_mul_d(mkdd(6.93147180559945286e-01, 2.31904681384629956e-17), (double)kadj));
  return dd_mul(l, mkdd(4.34294481903251817e-01, 1.09831965021676507e-17)).h;
}

// ------------------------------------------------------------ sin / cos ----
// sin and cos of a double as double-doubles (|x| < ~1e8 keeps full accuracy).
LSL_HDN void sincos_dd(double x, dd* sn, dd* cs) {
  const double P1 = 1.57079632679489656e+00, P2 = 6.12323399573676604e-17,
               P3 = -1.49738490485916983e-33;
  double fn = floor(x * 6.36619772367581382433e-01 + 0.5);
  dd r = dd_add(mkdd(x, 0.0), dd_neg(two_prod(fn, P1)));
  r = dd_add(r, dd_neg(two_prod(fn, P2)));
  r = dd_add_d(r, -(fn * P3));
  dd z = dd_mul(r, r);
  // sin r = r (1 - z/3! + z^2/5! - ... + z^13/27!)
  dd p = mkdd(-9.18368986379554601e-29, -1.43031503967873220e-45);          // -1/27!
  p = dd_add(dd_mul(p, z), mkdd(6.44695028438447359e-26, -1.93304042337034648e-42));   // 1/25!
  p = dd_add(dd_mul(p, z), mkdd(-3.86817017063068413e-23, 8.84317765548234385e-40));   // -1/23!
  p = dd_add(dd_mul(p, z), mkdd(1.95729410633912626e-20, -1.36435038300879085e-36));   // 1/21!
  p = dd_add(dd_mul(p, z), mkdd(-8.22063524662432950e-18, -2.21418941196042654e-34));  // -1/19!
  p = dd_add(dd_mul(p, z), mkdd(2.81145725434552060e-15, 1.65088427308614326e-31));    // 1/17!
  p = dd_add(dd_mul(p, z), mkdd(-7.64716373181981641e-13, -7.03872877733453001e-30));  // -1/15!
  p = dd_add(dd_mul(p, z), mkdd(1.60590438368216133e-10, 1.25852945887520981e-26));    // 1/13!
  p = dd_add(dd_mul(p, z), mkdd(-2.50521083854417202e-08, 1.44881407093591197e-24));   // -1/11!
  p = dd_add(dd_mul(p, z), mkdd(2.75573192239858925e-06, -1.85839327404647208e-22));   // 1/9!
  p = dd_add(dd_mul(p, z), mkdd(-1.98412698412698413e-04, -1.72095582934207053e-22));  // -1/7!
  p = dd_add(dd_mul(p, z), mkdd(8.33333333333333322e-03, 1.15648231731787138e-19));    // 1/5!
  p = dd_add(dd_mul(p, z), mkdd(-1.66666666666666657e-01, -9.25185853854297066e-18));  // -1/3!
  dd s = dd_add(r, dd_mul(dd_mul(p, z), r));
  // cos r = 1 - z/2! + z^2/4! - ... + z^14/28!
  dd q = mkdd(3.27988923706983776e-30, 1.51175427440298787e-46);            // 1/28!
  q = dd_add(dd_mul(q, z), mkdd(-2.47959626322479759e-27, 1.29537309647652288e-43));   // -1/26!
  q = dd_add(dd_mul(q, z), mkdd(1.61173757109611839e-24, -3.68465735645097660e-41));   // 1/24!
  q = dd_add(dd_mul(q, z), mkdd(-8.89679139245057408e-22, 7.91140261487237622e-38));   // -1/22!
  q = dd_add(dd_mul(q, z), mkdd(4.11031762331216484e-19, 1.44129733786595271e-36));    // 1/20!
  q = dd_add(dd_mul(q, z), mkdd(-1.56192069685862253e-16, -1.19106796602737540e-32));  // -1/18!
  q = dd_add(dd_mul(q, z), mkdd(4.77947733238738525e-14, 4.39920548583408126e-31));    // 1/16!
  q = dd_add(dd_mul(q, z), mkdd(-1.14707455977297245e-11, -2.06555127528307454e-28));  // -1/14!
  q = dd_add(dd_mul(q, z), mkdd(2.08767569878681002e-09, -1.20734505911325997e-25));   // 1/12!
  q = dd_add(dd_mul(q, z), mkdd(-2.75573192239858883e-07, -2.37677146222502973e-23));  // -1/10!
  q = dd_add(dd_mul(q, z), mkdd(2.48015873015873016e-05, 2.15119478667758816e-23));    // 1/8!
  q = dd_add(dd_mul(q, z), mkdd(-1.38888888888888894e-03, 5.30054395437357706e-20));   // -1/6!
  q = dd_add(dd_mul(q, z), mkdd(4.16666666666666644e-02, 2.31296463463574266e-18));    // 1/4!
  q = dd_add(dd_mul(q, z), mkdd(-0.5, 0.0));                                           // -1/2!
  dd c = dd_add_d(dd_mul(q, z), 1.0);
  switch (((int)fn) & 3) {
    case 0: *sn = s; *cs = c; break;
    case 1: *sn = c; *cs = dd_neg(s); break;
    case 2: *sn = dd_neg(s); *cs = dd_neg(c); break;
    default: *sn = dd_neg(c); *cs = s; break;
  }
}
LSL_HD double lsl_sin(double x) {
  if (x != x || x - x != 0.0) return LSL_NAN;
  if (fabs(x) < 7.4505805969238281e-09) return x;
  dd s, c; sincos_dd(x, &s, &c); return s.h;
}
LSL_HD double lsl_cos(double x) {
  if (x != x || x - x != 0.0) return LSL_NAN;
  if (fabs(x) < 7.4505805969238281e-09) return 1.0;
  dd s, c; sincos_dd(x, &s, &c); return c.h;
}
LSL_HD void lsl_sincos(double x, double* s, double* c) {
  if (x != x || x - x != 0.0) { *s = *c = LSL_NAN; return; }
  if (fabs(x) < 7.4505805969238281e-09) { *s = x; *c = 1.0; return; }
  dd a, b; sincos_dd(x, &a, &b); *s = a.h; *c = b.h;
}

// ---------------------------------------------------------- atan / atan2 ----
LSL_HD double lsl_atan2(double y, double x) {
  if (x != x || y != y) return x + y;
  int sy = (int)(d2u(y) >> 63), sx = (int)(d2u(x) >> 63);
  int m = sy | (sx << 1);
  if (y == 0.0) {
    switch (m) { case 0: case 1: return y; case 2: return LSL_PI; default: return -LSL_PI; }
  }
  if (x == 0.0) return sy ? -LSL_PIO2 : LSL_PIO2;
  double ax = fabs(x), ay = fabs(y);
  if (ax == HUGE_VAL) {
    if (ay == HUGE_VAL) {
      switch (m) { case 0: return 0.25 * LSL_PI; case 1: return -0.25 * LSL_PI;
                   case 2: return 0.75 * LSL_PI; default: return -0.75 * LSL_PI; }
    }
    switch (m) { case 0: return 0.0; case 1: return -0.0; case 2: return LSL_PI; default: return -LSL_PI; }
  }
  if (ay == HUGE_VAL) return sy ? -LSL_PIO2 : LSL_PIO2;
  double z0 = seed_atan2(y, x);
  int k = (int)((d2u(ay) >> 52) & 0x7ff) - (int)((d2u(ax) >> 52) & 0x7ff);
  if (k > 60 || k < -60) return z0;
  // atan2(y,x) - z0 = atan2(y cos z0 - x sin z0, x cos z0 + y sin z0) ~ the tiny ratio itself
  dd s, c; sincos_dd(z0, &s, &c);
  dd num = dd_add(dd_mul_d(c, y), dd_neg(dd_mul_d(s, x)));
  dd den = dd_add(dd_mul_d(c, x), dd_mul_d(s, y));
  double corr = num.h / den.h;
  return two_sum(z0, corr).h;
}
// acos for |x| <= 1 through the correctly rounded atan2 (sqrt is IEEE): used only for angle thresholds
// (src/line/motion.cpp:452). |x| > 1 (rounding) gives NaN like libm.
LSL_HD double lsl_acos(double x) { return lsl_atan2(sqrt((1.0 - x) * (1.0 + x)), x); }
LSL_HD double lsl_atan(double x) { return lsl_atan2(x, 1.0); }

// ----------------------------------------------------------- pow / sinh ----
LSL_HD double lsl_pow(double x, double y) {
  if (y == 0.0) return 1.0;
  if (x != x || y != y) return x + y;
  if (y == 1.0) return x;
  if (y == 2.0) return x * x;
  double yf = floor(y);
  int yint = (yf == y) && fabs(y) < 9007199254740992.0;
  int yodd = yint && (yf * 0.5 != floor(yf * 0.5));
  if (x == 0.0) { if (y > 0) return yodd ? x : 0.0; return HUGE_VAL; }
  double sgn = 1.0;
  if (x < 0.0) { if (!yint) return LSL_NAN; if (yodd) sgn = -1.0; x = -x; }
  if (x == HUGE_VAL) return y > 0 ? sgn * HUGE_VAL : 0.0;
  if (x == 1.0) return sgn;
  int kadj = 0;
  if ((d2u(x) >> 52) == 0) { x *= 18014398509481984.0; kadj = -54; }
  dd l = log_dd(x);
  if (kadj) l = dd_add(l, dd_mul_d(mkdd(6.93147180559945286e-01, 2.31904681384629956e-17), (double)kadj));
  dd p = dd_mul_d(l, y);
  if (p.h > 7.09782712893383973096e+02) return sgn * HUGE_VAL;
  if (p.h < -7.45133219101941108420e+02) return sgn * 0.0;
  int k; dd e = exp_dd(p, &k);
  return sgn * scale2(e.h, k);
}

LSL_HD double lsl_sinh(double x) {
  if (x != x) return x;
  double ax = fabs(x), r;
  if (ax < 7.4505805969238281e-09) return x;
  if (ax < 0.3) {
    dd xx = mkdd(ax, 0.0), z = two_prod(ax, ax);
    dd p = mkdd(6.44695028438447359e-26, -1.93304042337034648e-42);                    // 1/25!
    p = dd_add(dd_mul(p, z), mkdd(3.86817017063068413e-23, -8.84317765548234385e-40)); // 1/23!
    p = dd_add(dd_mul(p, z), mkdd(1.95729410633912626e-20, -1.36435038300879085e-36)); // 1/21!
    p = dd_add(dd_mul(p, z), mkdd(8.22063524662432950e-18, 2.21418941196042654e-34));  // 1/19!
    p = dd_add(dd_mul(p, z), mkdd(2.81145725434552060e-15, 1.65088427308614326e-31));  // 1/17!
    p = dd_add(dd_mul(p, z), mkdd(7.64716373181981641e-13, 7.03872877733453001e-30));  // 1/15!
    p = dd_add(dd_mul(p, z), mkdd(1.60590438368216133e-10, 1.25852945887520981e-26));  // 1/13!
    p = dd_add(dd_mul(p, z), mkdd(2.50521083854417202e-08, -1.44881407093591197e-24)); // 1/11!
    p = dd_add(dd_mul(p, z), mkdd(2.75573192239858925e-06, -1.85839327404647208e-22)); // 1/9!
    p = dd_add(dd_mul(p, z), mkdd(1.98412698412698413e-04, 1.72095582934207053e-22));  // 1/7!
    p = dd_add(dd_mul(p, z), mkdd(8.33333333333333322e-03, 1.15648231731787138e-19));  // 1/5!
    p = dd_add(dd_mul(p, z), mkdd(1.66666666666666657e-01, 9.25185853854297066e-18));  // 1/3!
    r = dd_add(xx, dd_mul(dd_mul(p, z), xx)).h;
  } else if (ax < 7.09e2) {
    int k; dd e = exp_dd(mkdd(ax, 0.0), &k);
    e.h = scale2(e.h, k); e.l = scale2(e.l, k);
    dd ei = dd_div(mkdd(1.0, 0.0), e);
    dd d = dd_add(e, dd_neg(ei));
    r = 0.5 * d.h;
  } else {
    int k; dd e = exp_dd(mkdd(ax, 0.0), &k);
    r = scale2(e.h, k - 1);
  }
  return x < 0 ? -r : r;
}

}  // namespace lslm
