/*
 * aff_detmath.h — deterministic FP64 elementary functions for host and device.
 *
 * The predictor thresholds accumulated FP64 values (< 1.0, > 0.05, < 0.001;
 * reference src/TrajectoryPredictor.cpp:304,768,832), so a verdict can flip on
 * a 1-ulp difference between glibc's and CUDA's libm.  These routines use only
 * IEEE +,-,*,/,sqrt and explicit fma(), all of which round identically on
 * x86-64 (FMA3) and sm_100a, so the CUDA path and its CPU checker see the same
 * bits.  Build rules that go with this header:
 *   device: nvcc -fmad=false            (no implicit contraction)
 *   host  : gcc  -ffp-contract=off -mfma (fma() -> one vfmadd instruction)
 *
 * Algorithms: Cody-Waite three-term reduction by pi/2 and the classic
 * minimax kernels for sin/cos on [-pi/4, pi/4] and atan with four breakpoints
 * (the published fdlibm coefficient sets).  Accuracy is about 1 ulp for the
 * argument ranges of this code base (|x| < 1e5 rad).
 */
#ifndef AFF_DETMATH_H
#define AFF_DETMATH_H

#include <math.h>

#if defined(__CUDACC__)
#define AFF_HD __host__ __device__ __forceinline__
#else
#define AFF_HD static inline
#endif

#define AFF_PI 3.14159265358979323846
#define AFF_PI_LO 1.2246467991473531772e-16
#define AFF_PI_2 1.57079632679489661923

/* sin and cos of r, |r| <= pi/4 (+ a little) */
AFF_HD double aff_ksin(double r)
{
  const double S1 = -1.66666666666666324348e-01, S2 = 8.33333333332248946124e-03,
               S3 = -1.98412698298579493134e-04, S4 = 2.75573137070700676789e-06,
               S5 = -2.50507602534068634195e-08, S6 = 1.58969099521155010221e-10;
  double z = r * r;
  double v = z * r;
  double p = fma(z, S6, S5);
  p = fma(z, p, S4);
  p = fma(z, p, S3);
  p = fma(z, p, S2);
  p = fma(z, p, S1);
  return fma(v, p, r);
}

AFF_HD double aff_kcos(double r)
{
  const double C1 = 4.16666666666666019037e-02, C2 = -1.38888888888741095749e-03,
               C3 = 2.48015872894767294178e-05, C4 = -2.75573143513906633035e-07,
               C5 = 2.08757232129817482790e-09, C6 = -1.13596475577881948265e-11;
  double z = r * r;
  double p = fma(z, C6, C5);
  p = fma(z, p, C4);
  p = fma(z, p, C3);
  p = fma(z, p, C2);
  p = fma(z, p, C1);
  double hz = 0.5 * z;
  double w = 1.0 - hz;
  return w + (((1.0 - w) - hz) + (z * z) * p);
}

AFF_HD void aff_sincos(double x, double* s, double* c)
{
  const double TWO_OVER_PI = 6.36619772367581382433e-01;
  const double P1 = 1.57079632673412561417e+00;  /* first 33 bits of pi/2 */
  const double P2 = 6.07710050630396597660e-11;  /* next 33 bits          */
  const double P3 = 2.02226624871116645580e-21;  /* next 33 bits          */
  const double MAGIC = 6755399441055744.0;       /* 1.5 * 2^52            */
  double fn = fma(x, TWO_OVER_PI, MAGIC) - MAGIC; /* nearest integer to 2x/pi */
  double r = fma(-fn, P1, x);
  r = fma(-fn, P2, r);
  r = fma(-fn, P3, r);
  long long n = (long long)fn;
  double ks = aff_ksin(r), kc = aff_kcos(r);
  switch ((int)(n & 3)) {
    case 0: *s = ks; *c = kc; break;
    case 1: *s = kc; *c = -ks; break;
    case 2: *s = -ks; *c = -kc; break;
    default: *s = -kc; *c = ks; break;
  }
}

AFF_HD double aff_sin(double x) { double s, c; aff_sincos(x, &s, &c); return s; }
AFF_HD double aff_cos(double x) { double s, c; aff_sincos(x, &s, &c); return c; }

/* atan for x >= 0 */
AFF_HD double aff_katan(double x)
{
  const double aT0 = 3.33333333333329318027e-01, aT1 = -1.99999999998764832476e-01,
               aT2 = 1.42857142725034663711e-01, aT3 = -1.11111104054623557880e-01,
               aT4 = 9.09088713343650656196e-02, aT5 = -7.69187620504482999495e-02,
               aT6 = 6.66107313738753120669e-02, aT7 = -5.83357013379057348645e-02,
               aT8 = 4.97687799461593236017e-02, aT9 = -3.65315727442169155270e-02,
               aT10 = 1.62858201153657823623e-02;
  double hi, lo;
  int direct = 0;
  if (x < 0.4375) {
    direct = 1; hi = 0.0; lo = 0.0;
  } else if (x < 0.6875) {
    hi = 4.63647609000806093515e-01; lo = 2.26987774529616870924e-17;
    x = (2.0 * x - 1.0) / (2.0 + x);
  } else if (x < 1.1875) {
    hi = 7.85398163397448278999e-01; lo = 3.06161699786838301793e-17;
    x = (x - 1.0) / (x + 1.0);
  } else if (x < 2.4375) {
    hi = 9.82793723247329054082e-01; lo = 1.39033110312309984516e-17;
    x = (x - 1.5) / (1.0 + 1.5 * x);
  } else {
    hi = 1.57079632679489655800e+00; lo = 6.12323399573676603587e-17;
    x = -1.0 / x;   /* x = +inf gives -0: result hi+lo */
  }
  double z = x * x;
  double w = z * z;
  double s1 = fma(w, aT10, aT8);
  s1 = fma(w, s1, aT6);
  s1 = fma(w, s1, aT4);
  s1 = fma(w, s1, aT2);
  s1 = fma(w, s1, aT0);
  s1 = z * s1;
  double s2 = fma(w, aT9, aT7);
  s2 = fma(w, s2, aT5);
  s2 = fma(w, s2, aT3);
  s2 = fma(w, s2, aT1);
  s2 = w * s2;
  if (direct) return x - x * (s1 + s2);
  return hi - ((x * (s1 + s2) - lo) - x);
}

AFF_HD double aff_atan2(double y, double x)
{
  if (y == 0.0) return (x < 0.0) ? AFF_PI : 0.0;
  if (x == 0.0) return (y > 0.0) ? AFF_PI_2 : -AFF_PI_2;
  double z = aff_katan(fabs(y / x));
  if (x > 0.0) return (y > 0.0) ? z : -z;
  return (y > 0.0) ? AFF_PI - (z - AFF_PI_LO) : (z - AFF_PI_LO) - AFF_PI;
}

/* acos with the argument clipped to [-1, 1] (Rcs Math_acos behaviour) */
AFF_HD double aff_acos(double x)
{
  if (x >= 1.0) return 0.0;
  if (x <= -1.0) return AFF_PI;
  return 2.0 * aff_atan2(sqrt(1.0 - x), sqrt(1.0 + x));
}

#endif /* AFF_DETMATH_H */
