// TEST / MEASUREMENT INFRASTRUCTURE, never linked into libaffpredict.so.
//
// Counts the FP64 operations of one rollout by running the KERNEL SOURCE (sweep.cuh, the
// thread-per-candidate kernel; same operations as the warp-per-candidate kernel) on the CPU with
// `double` replaced by a counting scalar.  The figure is what the algorithm of the product needs per
// rollout with the plan's static / dynamic split (static bodies are not recomputed, structurally zero
// Jacobian blocks are not formed): it is the `flops_per_rollout` of bench.py's FP64 roofline
// ("the counted figure is the one to report"), reproducible with tools/count_flops.py.
// FMA counts 2, add / sub / mul / div / sqrt 1 each; comparisons, negations, fabs and integer work 0.
#define AFF_EMULATE 1
#include <float.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

struct AffFlopCounters { unsigned long long add, mul, fma, div, sqrt; };
static AffFlopCounters g_cnt;

struct CD
{
  double v;
  CD() = default;
  CD(double x) : v(x) {}
  explicit operator double() const { return v; }
  explicit operator float() const { return (float)v; }
  explicit operator long long() const { return (long long)v; }
  explicit operator int() const { return (int)v; }
  explicit operator bool() const { return v != 0.0; }
  CD& operator+=(CD o) { ++g_cnt.add; v += o.v; return *this; }
  CD& operator-=(CD o) { ++g_cnt.add; v -= o.v; return *this; }
  CD& operator*=(CD o) { ++g_cnt.mul; v *= o.v; return *this; }
  CD& operator/=(CD o) { ++g_cnt.div; v /= o.v; return *this; }
};
template <class T> using Arith = typename std::enable_if<std::is_arithmetic<T>::value, int>::type;
#define CD_BINOP(op, counter)                                                                        \
  static inline CD operator op(CD a, CD b) { ++g_cnt.counter; return CD(a.v op b.v); }               \
  template <class T, Arith<T> = 0> static inline CD operator op(CD a, T b) { ++g_cnt.counter; return CD(a.v op (double)b); } \
  template <class T, Arith<T> = 0> static inline CD operator op(T a, CD b) { ++g_cnt.counter; return CD((double)a op b.v); }
CD_BINOP(+, add)
CD_BINOP(-, add)
CD_BINOP(*, mul)
CD_BINOP(/, div)
#define CD_CMP(op)                                                                                   \
  static inline bool operator op(CD a, CD b) { return a.v op b.v; }                                  \
  template <class T, Arith<T> = 0> static inline bool operator op(CD a, T b) { return a.v op (double)b; } \
  template <class T, Arith<T> = 0> static inline bool operator op(T a, CD b) { return (double)a op b.v; }
CD_CMP(<)
CD_CMP(>)
CD_CMP(<=)
CD_CMP(>=)
CD_CMP(==)
CD_CMP(!=)
static inline CD operator-(CD a) { return CD(-a.v); }
static inline CD operator+(CD a) { return a; }
static inline CD fma(CD a, CD b, CD c) { ++g_cnt.fma; return CD(::fma(a.v, b.v, c.v)); }
static inline CD sqrt(CD a) { ++g_cnt.sqrt; return CD(::sqrt(a.v)); }
static inline CD fabs(CD a) { return CD(::fabs(a.v)); }

#define double CD
#include "aff_detmath.h"
#include "kernel/plan.h"
#include "kernel/warp.cuh"
// the warp primitives are only referenced by the warp-per-candidate kernel, which is not run here
// (the int versions and w_lane / w_sync / w_ballot are defined by warp_emu.cpp in the same library)
CD w_shfl(CD v, int) { return v; }
CD w_shfl_xor(CD v, int) { return v; }
#include "kernel/rollout.cuh"
#include "kernel/sweep.cuh"
#undef double

// `plan` is a KPlan built (with plain doubles: same layout) by the emulator's driver, prologue already run.
extern "C" void aff_flop_count_rollout(const void* plan, int cand, unsigned long long out[5])
{
  static_assert(sizeof(CD) == sizeof(double), "counting scalar must have the layout of double");
  const KPlan& P = *(const KPlan*)plan;
  TState* S = (TState*)calloc(1, sizeof(TState));
  memset(&g_cnt, 0, sizeof(g_cnt));
  t_rollout(P, cand, *S, true);
  free(S);
  out[0] = g_cnt.add; out[1] = g_cnt.mul; out[2] = g_cnt.fma; out[3] = g_cnt.div; out[4] = g_cnt.sqrt;
}
