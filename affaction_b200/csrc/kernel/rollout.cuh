// One candidate rollout on one warp (or on one block: see below): the whole of
//   TrajectoryPredictor::predict     (reference src/TrajectoryPredictor.cpp:79-406)
//   TrajectoryPredictor::computeIK   (:573-753)
//   TrajectoryPredictor::checkState  (:756-886)
// fused into a single step loop whose state lives in shared memory.
//
// Lane mapping per phase (32 lanes):
//   trajectories ...... one lane per task dimension (1-D via-point polynomial)
//   task values / dx .. one lane per task
//   Jacobians, dH, dq . one lane per IK joint (column of J)
//   J W^-1 J^T ........ one lane per matrix entry; Cholesky one lane per row
//   forward kinematics  12 lanes per node (one per HTr element), two nodes per round
//   collision model ... one lane per shape / body pair
// Block-per-candidate launches (batches of a handful of candidates: k_rollout_t<1>, k_cta_* below) run the
// same step loop on warp 0 of a block and move what is not on the chain FK -> task values -> Jacobian ->
// solve -> FK to helper warps (collision model, trajectories, null-space terms and checks, orientation-task
// values): same expressions, same values, a third of the time per step.
// Every floating-point expression is written in the order the CPU checker uses
// (explicit fma(), compiled with -fmad=false), and the few reductions over
// lanes are order-independent (min / max / ballot) or use the fixed butterfly
// order of w_tree_sum(), so results are reproducible bit for bit.
//
// This file may be included twice, inside two namespaces: once as is (every shared-memory offset
// and matrix stride is read from the plan) and once with AFFK_FIXED_LAYOUT defined (offsets and
// strides of plan.h's fixed layout are compile-time constants: no address arithmetic per access).
// The host compiler of plans uses the fixed layout whenever a batch fits its caps, so both kernels
// see the same numbers in KPlan.
#include <float.h>
#include <math.h>

#include "aff_detmath.h"
#include "plan.h"
#include "warp.cuh"

#undef AFFK_O
#undef AFFK_JS
#undef AFFK_LS
#if defined(AFFK_FIXED_LAYOUT)
#define AFFK_O(P, name) ((int)AFFK_FIX_o_##name)
#define AFFK_JS(P) ((int)AFFK_FIX_JS)
#define AFFK_LS(P) ((int)AFFK_FIX_LS)
#else
#define AFFK_O(P, name) ((P).o_##name)
#define AFFK_JS(P) ((P).JS)
#define AFFK_LS(P) ((P).LS)
#endif

// -DAFFK_PROFILE (tuning builds only): lane 0 of warp 0 of block 0 adds up the clock cycles of the parts of a step
#if defined(AFFK_PROFILE) && !defined(AFF_EMULATE)
__device__ unsigned long long g_kClocks[16];
#define K_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = clock64(); g_kClocks[i] += (unsigned long long)(now_ - kmark_); kmark_ = now_; } } while (0)
#define K_MARK_INIT long long kmark_ = clock64()
#define K_SUB(i, t0_) do { if (blockIdx.x == 0 && threadIdx.x == 0) { g_kClocks[i] += (unsigned long long)(clock64() - (t0_)); } } while (0)
#define K_NOW() clock64()
#else
#define K_MARK(i) do { } while (0)
#define K_MARK_INIT do { } while (0)
#define K_SUB(i, t0_) do { } while (0)
#define K_NOW() 0ll
#endif
#define AFFK_TRACK_MAX 0.05      // MAX_TRACKING_ERROR, src/TrajectoryPredictor.cpp:50
#define AFFK_TRAJ_EPS 1.0e-8
#define AFFK_NO_LIMIT 1.0e300
#define AFFK_GOLDEN_ITERS 48

// ------------------------------------------------------------ small algebra
AFF_DEV void k_cross(double* o, const double* a, const double* b)
{
  o[0] = fma(a[1], b[2], -(a[2] * b[1]));
  o[1] = fma(a[2], b[0], -(a[0] * b[2]));
  o[2] = fma(a[0], b[1], -(a[1] * b[0]));
}
AFF_DEV double k_dot(const double* a, const double* b) { return fma(a[2], b[2], fma(a[1], b[1], a[0] * b[0])); }
// o = R v, R = 9 doubles row-major
AFF_DEV void k_mulv(double* o, const double* R, const double* v)
{
  for (int i = 0; i < 3; ++i) o[i] = fma(R[3 * i + 2], v[2], fma(R[3 * i + 1], v[1], R[3 * i] * v[0]));
}
AFF_DEV double k_clip(double v, double lo, double hi) { return v < lo ? lo : (v > hi ? hi : v); }
AFF_DEV double k_clamp01(double v) { return v < 0.0 ? 0.0 : (v > 1.0 ? 1.0 : v); }

// element e (0..2 org, 3..11 rot) of T o P
AFF_DEV double k_compose_elem(const double* T, const double* P, int e)
{
  if (e < 3) return fma(P[3 + 6 + e], T[2], fma(P[3 + 3 + e], T[1], fma(P[3 + e], T[0], P[e])));
  const int i = (e - 3) / 3, j = (e - 3) % 3;
  return fma(T[3 + 3 * i + 2], P[3 + 6 + j], fma(T[3 + 3 * i + 1], P[3 + 3 + j], T[3 + 3 * i] * P[3 + j]));
}

// Serial versions used for rare per-node work (object nodes, prologue).
AFF_DEV void k_compose(double* out, const double* T, const double* P)
{
  double r[12];
  for (int e = 0; e < 12; ++e) r[e] = k_compose_elem(T, P, e);
  for (int e = 0; e < 12; ++e) out[e] = r[e];
}
AFF_DEV void k_apply_joint(double* A, int type, double q)
{
  if (type < AFF_JNT_ROT_X) {
    for (int k = 0; k < 3; ++k) A[k] = fma(q, A[3 + 3 * type + k], A[k]);
  } else {
    const int d = type - AFF_JNT_ROT_X, a = (d + 1) % 3, b = (d + 2) % 3;
    double s, c;
    aff_sincos(q, &s, &c);
    for (int k = 0; k < 3; ++k) {
      const double ra = A[3 + 3 * a + k], rb = A[3 + 3 * b + k];
      A[3 + 3 * a + k] = fma(s, rb, c * ra);
      A[3 + 3 * b + k] = fma(c, rb, -(s * ra));
    }
  }
}

// ------------------------------------------------------------------ context
// Every body frame the step loop touches lives in shared memory as a "slot":
// slots [0, nDyn) are the dynamic nodes, the following ones are copies of the
// static frames the loop refers to (parents of dynamic nodes, task reference
// bodies, base_footprint ...), made once per rollout.  -1 means "none / world".
struct KCtx
{
  const KPlan* P;
  const KCand* C;
  double* sm;        // warp base in shared memory
  double* nodes;     // [nSlots][12]
  unsigned* chain;   // [nSlots] IK joints (bit = jacobi index) that currently move the slot
  int* tdesc;        // [nTasks][8] kind, eff, ref, frame, axis, xoff, dim, n_joints
  int lane;
  // re-parentable objects (warp-uniform)
  // (two explicit slots instead of arrays: a dynamically indexed member array would push the whole
  //  context, pointers included, into local memory)
  int objParentSlot0, objParentSlot1;
  int objParentDynamic0, objParentDynamic1;
  unsigned objChain0, objChain1;
  int objTree0, objTree1;
  int objMode0, objMode1;       // 0 per-shape flags, 1 all on, 2 all off
  unsigned objHasRel;
  // accessors shared with the thread-per-candidate context of sweep.cuh
  AFF_MEMBER const double* qv() const { return sm + AFFK_O(*P, q); }
  AFF_MEMBER const double* shpv() const { return sm + AFFK_O(*P, shp); }
};
#if AFFK_MAX_OBJ != 2
#error "object state is written out for two slots"
#endif
#define OBJ_GET(c, f, s) ((s) ? (c).f##1 : (c).f##0)
#define OBJ_SET(c, f, s, v) do { if (s) (c).f##1 = (v); else (c).f##0 = (v); } while (0)

template <class Ctx> AFF_DEV const double* k_slot(const Ctx& c, int slot) { return c.nodes + 12 * slot; }

// Jacobian columns of this lane's IK joint for a point p fixed to the body in `slot`.
struct KJointFrame { double ax[3]; double org[3]; int type; int valid; };

template <class Ctx> AFF_DEV void k_jac_cols(const Ctx& c, const KJointFrame& jf, int col, int slot, const double* p, double* colT, double* colR)
{
  colT[0] = colT[1] = colT[2] = 0.0;
  colR[0] = colR[1] = colR[2] = 0.0;
  if (slot < 0 || !jf.valid || !((c.chain[slot] >> col) & 1u)) return;
  if (jf.type < AFF_JNT_ROT_X) { colT[0] = jf.ax[0]; colT[1] = jf.ax[1]; colT[2] = jf.ax[2]; }
  else {
    double r[3];
    for (int k = 0; k < 3; ++k) r[k] = p[k] - jf.org[k];
    k_cross(colT, jf.ax, r);
    colR[0] = jf.ax[0]; colR[1] = jf.ax[1]; colR[2] = jf.ax[2];
  }
}

// ------------------------------------------------------- forward kinematics
// Pose of an object node that still hangs on its six rigid-body joints.
AFF_DEV void k_object_from_q6(const KCtx& c, int n, int slot)
{
  double A[12];
  const double* Pp = k_slot(c, OBJ_GET(c, objParentSlot, slot));
  for (int e = 0; e < 12; ++e) A[e] = Pp[e];
  const double* q6 = c.sm + c.P->o_objq6 + 6 * slot;
  for (int k = 0; k < 6; ++k) k_apply_joint(A, k, q6[k]);
  double* out = c.nodes + 12 * n;
  for (int e = 0; e < 12; ++e) out[e] = A[e];
}

// FK table entry, one per (round, lane): everything lane `l` needs to produce its HTr element of
// the node scheduled in its group this round (plan_build.cpp: buildFkTable).
//   w0: bits 0-11 parent node base (doubles from the node array; 0 unless the lane composes), 12-23
//       destination element (idle lanes: the spare element behind the array), 24-26 op, 27-31 partner lane
//   w1: bits 0-5 jacobi index (IK joints; object slot for objects) | bit 7 constant joint | bits 8-15
//       dynamic node (constant joints, objects) | bits 16-27 shared-memory index of the lane's extra
//       operand: the parent's element (COPY), q or sin(q) of the joint (IK joints), else 0
//   t0..t2: the three entries of the node's fixed transform this lane multiplies with
// Element e of T o P:  org: P[e] + T0 P[3+e] + T1 P[6+e] + T2 P[9+e];  rot(i,j): T[i][0] P[3+j] + T[i][1] P[6+j] + T[i][2] P[9+j]
#define FKOP_IDLE 0
#define FKOP_COPY 1      /* identity transform: copy the parent's element            */
#define FKOP_FIXED 2     /* compose only                                            */
#define FKOP_ROT_A 3     /* compose, then ra' = c ra + s rb  (partner holds rb)      */
#define FKOP_ROT_B 4     /* compose, then rb' = c rb - s ra  (partner holds ra)      */
#define FKOP_TRANS 5     /* compose, then org += q * axis    (partner holds axis)    */
#define FKOP_OBJECT 6    /* re-parentable object: transform and parent are run-time  */

// RcsGraph_setState for the dynamic nodes: 12 lanes per node (one per HTr element), two nodes
// per round on the host-built schedule; all parents are shared-memory slots.
// rounds [r0, r1) of the schedule; r0 == 0 also refreshes sin / cos of the IK joints
template <int PREFETCH> AFF_DEV void k_fk_t(const KCtx& c, bool init, int r0, int r1)
{
  const KPlan& P = *c.P;
  double* sn = c.sm + AFFK_O(P, sn); double* cs = c.sm + AFFK_O(P, cs);
  const double* q = c.sm + AFFK_O(P, q);
  double* nodes = c.nodes;
  const int lane = c.lane;
  const long long tfk_ = K_NOW();
  if (r0 == 0) {
    if (lane < P.nJ) { double s, co; aff_sincos(q[lane], &s, &co); sn[lane] = s; cs[lane] = co; }
    w_sync();
  }
  K_SUB(11, tfk_);
  const int e = lane % 12;
  const int col = (e < 3) ? e : (e - 3) % 3, eo = (e < 3) ? e : 0;
  const KFkEntry* tab = P.fktab + lane + 32 * r0;
  // bit 0 = this round
  unsigned constRounds = (r0 < 32) ? P.fk_const_rounds[0] >> r0 : P.fk_const_rounds[1] >> (r0 - 32);
  unsigned objRounds = (r0 < 32) ? P.fk_obj_rounds[0] >> r0 : P.fk_obj_rounds[1] >> (r0 - 32);
  const int scratchDst = 12 * (P.nDyn + P.nStatPar);      // one spare element behind the node array
  // PREFETCH (block-per-candidate kernel: registers to spare): the entry of the next round is in flight
  // while this round computes (the table has one spare round behind its end)
  KFkEntry nxt;
  if (PREFETCH) nxt = ldg_fk(tab);
  for (int r = r0; r < r1; ++r, tab += 32) {
    KFkEntry en;
    if (PREFETCH) { en = nxt; nxt = ldg_fk(tab + 32); }
    else en = ldg_fk(tab);
    const int op = (int)((en.w0 >> 24) & 7u);
    const int pbase = (int)(en.w0 & 4095u);               // parent node (0 for lanes that do not compose)
    int dst = (int)((en.w0 >> 12) & 4095u);                // idle lanes: the spare element
    const int partner = (int)(en.w0 >> 27);
    const int aux = (int)((en.w1 >> 16) & 4095u);          // COPY: the parent's element; joints: q / sin entry; else 0
    const int qi = (int)(en.w1 & 63u);
    // The common operations are straight-line code: every lane reads its operands from addresses the
    // table made valid for it, computes all variants and selects; objects and constant joints under a
    // moving parent sit in rounds the plan has flagged.
    const double* Pp = nodes + pbase;
    const double p0 = Pp[eo], p3 = Pp[3 + col], p6 = Pp[6 + col], p9 = Pp[9 + col];
    const double x = c.sm[aux];
    const double acc0 = (e < 3) ? p0 : 0.0;
    const double comp = fma(en.t2, p9, fma(en.t1, p6, fma(en.t0, p3, acc0)));
    const bool compose = (op >= FKOP_FIXED) && (op <= FKOP_TRANS);
    double val = (op == FKOP_COPY) ? x : (compose ? comp : 0.0);
    double s = x, co = cs[qi];
    if (r == 32) { constRounds = P.fk_const_rounds[1]; objRounds = P.fk_obj_rounds[1]; }
    if (constRounds & 1u) {
      if (op > FKOP_FIXED && op <= FKOP_TRANS && (en.w1 & 0x80u)) {
        const int n = (int)((en.w1 >> 8) & 255u);
        s = (op == FKOP_TRANS) ? ldg(&P.dynQ[n]) : ldg(&P.dynSC[2 * n]);
        co = ldg(&P.dynSC[2 * n + 1]);
      }
    }
    if (objRounds & 1u) {
      if (op == FKOP_OBJECT) {
        const int slot = (int)(en.w1 & 3u), n = (int)((en.w1 >> 8) & 255u);
        if ((c.objHasRel >> slot) & 1u) {
          const double* Po = nodes + 12 * OBJ_GET(c, objParentSlot, slot);
          const double* R = c.sm + AFFK_O(P, objrel) + 12 * slot;
          const int tb = (e < 3) ? 0 : 3 + 3 * ((e - 3) / 3);
          const double a0 = (e < 3) ? Po[e] : 0.0;
          val = fma(R[tb + 2], Po[9 + col], fma(R[tb + 1], Po[6 + col], fma(R[tb], Po[3 + col], a0)));
        } else {
          dst = scratchDst;
          if ((init || OBJ_GET(c, objParentDynamic, slot)) && e == 0) k_object_from_q6(c, n, slot);
        }
      }
    }
    const double other = w_shfl(val, partner);
    const double rotA = fma(s, other, co * val), rotB = fma(co, val, -(s * other)), trn = fma(s, other, val);
    val = (op == FKOP_ROT_A) ? rotA : ((op == FKOP_ROT_B) ? rotB : ((op == FKOP_TRANS) ? trn : val));
    nodes[dst] = val;
    constRounds >>= 1; objRounds >>= 1;
    w_sync();
  }
}

AFF_DEV void k_fk(const KCtx& c, bool init) { k_fk_t<0>(c, init, 0, c.P->nRounds); }

// ------------------------------------------------------------ via-point step
// One receding-horizon step of a 1-D via-point trajectory (the oracle's
// oracle_via_step): window selection, reduced polynomial system, evaluation.
AFF_DEV void k_via_step(double* st, const KVia* vias, int n, double tnow, double dt, double* A, int maxUnk)
{
  if (n <= 0) { st[1] = 0.0; st[2] = 0.0; return; }
  int G = -1, lastFull = -1;
  for (int i = 0; i < n; ++i) {
    if (ldg(&vias[i].flag) == 7) { lastFull = i; if (ldg(&vias[i].t) - tnow >= AFFK_HORIZON) { G = i; break; } }
  }
  if (G < 0) G = (lastFull >= 0) ? lastFull : n - 1;
  int K = 0, nwin = 0;
  for (int i = 0; i <= G; ++i) {
    const int flag = ldg(&vias[i].flag);
    const int add = (flag & 1) + ((flag >> 1) & 1) + ((flag >> 2) & 1);
    if (K + add > maxUnk) break;
    K += add; nwin = i + 1;
  }
  if (K == 0) { st[1] = 0.0; st[2] = 0.0; return; }
  const double T = ldg(&vias[nwin - 1].t) - tnow;
  const double c0 = st[0], c1 = st[1] * T, c2 = (0.5 * st[2]) * (T * T);
  const int S = K + 1;
  double* sol = A + maxUnk * (maxUnk + 1);   // maxUnk doubles after the matrix
  if (K == 3 && nwin == 1) {
    // The common window: one full constraint, at tau = (t - tnow) / T = 1 exactly.  The system matrix
    // is then the constant {1 1 1; 3 4 5; 6 12 20}: the same elimination as below, written on a local
    // array so that the compiler folds the pivoting and the multipliers; only the right-hand side is
    // computed at run time, with the operations and their order unchanged.
    const double tau = 1.0, pw1 = 1.0 * tau, pw2 = pw1 * tau;
    double M[3][4];
    { double p = pw2; for (int k = 0; k < 3; ++k) { p = p * tau; M[0][k] = p; } }
    { double p = pw1; for (int k = 0; k < 3; ++k) { p = p * tau; M[1][k] = (double)(k + 3) * p; } }
    { double p = 1.0; for (int k = 0; k < 3; ++k) { p = p * tau; M[2][k] = (double)((k + 3) * (k + 2)) * p; } }
    M[0][3] = ldg(&vias[0].x) - fma(c2, pw2, fma(c1, tau, c0));
    M[1][3] = ldg(&vias[0].xd) * T - fma(2.0 * c2, tau, c1);
    M[2][3] = ldg(&vias[0].xdd) * (T * T) - 2.0 * c2;
#pragma unroll
    for (int cI = 0; cI < 3; ++cI) {
      int piv = cI;
      double best = fabs(M[cI][cI]);
#pragma unroll
      for (int r = cI + 1; r < 3; ++r) { const double v = fabs(M[r][cI]); if (v > best) { best = v; piv = r; } }
#pragma unroll
      for (int r = cI + 1; r < 3; ++r)
        if (piv == r) {
#pragma unroll
          for (int k = cI; k < 4; ++k) { const double tmp = M[cI][k]; M[cI][k] = M[r][k]; M[r][k] = tmp; }
        }
      const double p = M[cI][cI];
#pragma unroll
      for (int r = cI + 1; r < 3; ++r) {
        const double f = M[r][cI] / p;
#pragma unroll
        for (int k = cI + 1; k < 4; ++k) M[r][k] = fma(-f, M[cI][k], M[r][k]);
      }
    }
    double sl[3];
#pragma unroll
    for (int r = 2; r >= 0; --r) {
      double sacc = M[r][3];
#pragma unroll
      for (int k = r + 1; k < 3; ++k) sacc = fma(-M[r][k], sl[k], sacc);
      sl[r] = sacc / M[r][r];
    }
    sol[0] = sl[0]; sol[1] = sl[1]; sol[2] = sl[2];
  } else {
  int row = 0;
  for (int i = 0; i < nwin; ++i) {
    const int flag = ldg(&vias[i].flag);
    const double tau = (ldg(&vias[i].t) - tnow) / T;
    const double pw1 = 1.0 * tau, pw2 = pw1 * tau;
    if (flag & 1) {
      double p = pw2;
      for (int k = 0; k < K; ++k) { p = p * tau; A[row * S + k] = p; }
      A[row * S + K] = ldg(&vias[i].x) - fma(c2, pw2, fma(c1, tau, c0));
      ++row;
    }
    if (flag & 2) {
      double p = pw1;
      for (int k = 0; k < K; ++k) { p = p * tau; A[row * S + k] = (double)(k + 3) * p; }
      A[row * S + K] = ldg(&vias[i].xd) * T - fma(2.0 * c2, tau, c1);
      ++row;
    }
    if (flag & 4) {
      double p = 1.0;
      for (int k = 0; k < K; ++k) { p = p * tau; A[row * S + k] = (double)((k + 3) * (k + 2)) * p; }
      A[row * S + K] = ldg(&vias[i].xdd) * (T * T) - 2.0 * c2;
      ++row;
    }
  }
  // Gaussian elimination, partial pivoting, first maximal pivot
  for (int cI = 0; cI < K; ++cI) {
    int piv = cI;
    double best = fabs(A[cI * S + cI]);
    for (int r = cI + 1; r < K; ++r) { const double v = fabs(A[r * S + cI]); if (v > best) { best = v; piv = r; } }
    if (piv != cI) for (int k = cI; k < S; ++k) { const double tmp = A[cI * S + k]; A[cI * S + k] = A[piv * S + k]; A[piv * S + k] = tmp; }
    const double p = A[cI * S + cI];
    if (p == 0.0) continue;
    for (int r = cI + 1; r < K; ++r) {
      const double f = A[r * S + cI] / p;
      for (int k = cI + 1; k < S; ++k) A[r * S + k] = fma(-f, A[cI * S + k], A[r * S + k]);
    }
  }
  for (int r = K - 1; r >= 0; --r) {
    double sacc = A[r * S + K];
    for (int k = r + 1; k < K; ++k) sacc = fma(-A[r * S + k], sol[k], sacc);
    const double p = A[r * S + r];
    sol[r] = (p != 0.0) ? sacc / p : 0.0;
  }
  }
  const double s = (dt < T) ? dt / T : 1.0;
  double p = 0.0, pd = 0.0, pdd = 0.0;
  for (int k = K - 1; k >= 0; --k) {
    const int deg = k + 3;
    p = fma(p, s, sol[k]);
    pd = fma(pd, s, (double)deg * sol[k]);
    pdd = fma(pdd, s, (double)(deg * (deg - 1)) * sol[k]);
  }
  const double s2 = s * s, s3 = s2 * s;
  st[0] = fma(p, s3, fma(c2, s2, fma(c1, s, c0)));
  st[1] = fma(pd, s2, fma(2.0 * c2, s, c1)) / T;
  st[2] = fma(pdd, s, 2.0 * c2) / (T * T);
}

// ----------------------------------------------------------- task geometry
// Per task t (one lane each), from the current node transforms:
//   geo[t][0..11] kind-specific vectors, xcur[xoff..] = Task::computeX.
// XYZ/X/Y/Z : 0-2 pe, 3-5 pr, 6-8 r12
// POLAR     : 0-2 a (axis in ref frame), 3-5 e1, 6-8 e2, 9-11 pe
// INCLINATION: 0-2 u, 3-5 v, 6-8 n (unit, or zero), 9-11 pe
#define AFFK_GEO 12
// task descriptor words in shared memory
#define TD_KIND 0
#define TD_EFF 1
#define TD_REF 2
#define TD_FRAME 3
#define TD_AXIS 4
#define TD_XOFF 5
#define TD_DIM 6
#define TD_NJ 7

// (selects instead of t[k]: a dynamically indexed local array would live in local memory)
AFF_DEV double k_sel3(const double* v, int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }

AFF_DEV void k_tangent_basis(const double* a, double* e1, double* e2)
{
  int k = 0;
  double m = fabs(a[0]);
  if (fabs(a[1]) < m) { k = 1; m = fabs(a[1]); }
  if (fabs(a[2]) < m) { k = 2; }
  const double ak = k_sel3(a, k);
  double t[3];
  for (int i = 0; i < 3; ++i) { const double ti = -(ak * a[i]); t[i] = (i == k) ? ti + 1.0 : ti; }
  const double n = sqrt(k_dot(t, t));
  for (int i = 0; i < 3; ++i) e1[i] = t[i] / n;
  k_cross(e2, a, e1);
}

template <class Ctx> AFF_DEV void k_task_geo(const Ctx& c, const int* td, const KTask* tk, double* geo, double* xcur)
{
  const int kind = td[TD_KIND], eff = td[TD_EFF], ref = td[TD_REF], frame = td[TD_FRAME], axis = td[TD_AXIS];
  const int xoff = td[TD_XOFF];
  if (kind <= AFF_TASK_Z) {
    const double* Ae = k_slot(c, eff);
    double r[3], xr[3];
    for (int k = 0; k < 3; ++k) {
      geo[k] = Ae[k];
      const double prk = (ref >= 0) ? k_slot(c, ref)[k] : 0.0;
      geo[3 + k] = prk;
      r[k] = (ref >= 0) ? Ae[k] - prk : Ae[k];
      geo[6 + k] = r[k];
    }
    if (frame >= 0) k_mulv(xr, k_slot(c, frame) + 3, r);
    else { xr[0] = r[0]; xr[1] = r[1]; xr[2] = r[2]; }
    if (kind == AFF_TASK_XYZ) { xcur[xoff] = xr[0]; xcur[xoff + 1] = xr[1]; xcur[xoff + 2] = xr[2]; }
    else xcur[xoff] = k_sel3(xr, kind - AFF_TASK_X);
  } else if (kind == AFF_TASK_POLAR) {
    const double* Ae = k_slot(c, eff);
    const double* aI = Ae + 3 + 3 * axis;
    double a[3], e1[3], e2[3];
    if (ref >= 0) k_mulv(a, k_slot(c, ref) + 3, aI);
    else { a[0] = aI[0]; a[1] = aI[1]; a[2] = aI[2]; }
    k_tangent_basis(a, e1, e2);
    for (int k = 0; k < 3; ++k) { geo[k] = a[k]; geo[3 + k] = e1[k]; geo[6 + k] = e2[k]; geo[9 + k] = Ae[k]; }
    xcur[xoff] = aff_acos(a[2]);
    xcur[xoff + 1] = aff_atan2(a[1], a[0]);
  } else if (kind == AFF_TASK_INCLINATION) {
    const double* Ae = k_slot(c, eff);
    const double* aI = Ae + 3 + 3 * axis;
    double u[3], v[3], n[3];
    for (int k = 0; k < 3; ++k) u[k] = aI[k];
    if (ref >= 0) { const double* z = k_slot(c, ref) + 3 + 6; v[0] = z[0]; v[1] = z[1]; v[2] = z[2]; }
    else { v[0] = 0.0; v[1] = 0.0; v[2] = 1.0; }
    k_cross(n, v, u);
    const double sn = sqrt(k_dot(n, n));
    if (sn > 1.0e-8) { for (int k = 0; k < 3; ++k) n[k] = n[k] / sn; }
    else { n[0] = n[1] = n[2] = 0.0; }
    for (int k = 0; k < 3; ++k) { geo[k] = u[k]; geo[3 + k] = v[k]; geo[6 + k] = n[k]; geo[9 + k] = Ae[k]; }
    xcur[xoff] = aff_acos(k_dot(u, v));
  } else {   // JOINTS
    const int nj = td[TD_NJ];
    const double* q = c.qv();
    for (int i = 0; i < nj; ++i) {
      const int ji = ldg(&tk->jidx[i]);
      xcur[xoff + i] = (ji >= 0) ? q[ji] : ldg(&tk->jconst[i]);
    }
  }
}

// The two orientation kinds of k_task_geo for a whole warp at once (geometry warp of the block-per-candidate
// kernel): lanes gather their vectors, then every lane runs ONE normalisation (sqrt, three divisions) and ONE
// acos together instead of one branch after the other; expression by expression the same as k_task_geo.
// on: this lane has a POLAR or Inclination task (td valid).
template <class Ctx> AFF_DEV void k_task_geo_orient(const Ctx& c, const int* td, double* geo, double* xcur, bool on)
{
  const int kind = on ? td[TD_KIND] : -1;
  const bool polar = kind == AFF_TASK_POLAR, incl = kind == AFF_TASK_INCLINATION;
  const int eff = on ? td[TD_EFF] : 0, ref = on ? td[TD_REF] : -1, axis = on ? td[TD_AXIS] : 0, xoff = on ? td[TD_XOFF] : 0;
  const double* Ae = k_slot(c, eff);
  const double* aI = Ae + 3 + 3 * axis;
  double a[3] = {0.0, 0.0, 1.0}, v[3] = {0.0, 0.0, 1.0}, w[3] = {1.0, 0.0, 0.0};
  double acArg = 0.0;
  if (polar) {
    if (ref >= 0) k_mulv(a, k_slot(c, ref) + 3, aI);
    else { a[0] = aI[0]; a[1] = aI[1]; a[2] = aI[2]; }
    // k_tangent_basis up to its normalisation
    int k = 0;
    double m = fabs(a[0]);
    if (fabs(a[1]) < m) { k = 1; m = fabs(a[1]); }
    if (fabs(a[2]) < m) { k = 2; }
    const double ak = k_sel3(a, k);
    for (int i = 0; i < 3; ++i) { const double ti = -(ak * a[i]); w[i] = (i == k) ? ti + 1.0 : ti; }
    acArg = a[2];
  } else if (incl) {
    for (int i = 0; i < 3; ++i) a[i] = aI[i];
    if (ref >= 0) { const double* z = k_slot(c, ref) + 3 + 6; v[0] = z[0]; v[1] = z[1]; v[2] = z[2]; }
    k_cross(w, v, a);
    acArg = k_dot(a, v);
  }
  const double nn = sqrt(k_dot(w, w));
  const double w0 = w[0] / nn, w1 = w[1] / nn, w2 = w[2] / nn;
  const double ac = aff_acos(acArg);
  if (polar) {
    const double e1[3] = {w0, w1, w2};
    double e2[3];
    k_cross(e2, a, e1);
    for (int i = 0; i < 3; ++i) { geo[i] = a[i]; geo[3 + i] = e1[i]; geo[6 + i] = e2[i]; geo[9 + i] = Ae[i]; }
    xcur[xoff] = ac;
    xcur[xoff + 1] = aff_atan2(a[1], a[0]);
  } else if (incl) {
    const bool ok = nn > 1.0e-8;
    const double n[3] = {ok ? w0 : 0.0, ok ? w1 : 0.0, ok ? w2 : 0.0};
    for (int i = 0; i < 3; ++i) { geo[i] = a[i]; geo[3 + i] = v[i]; geo[6 + i] = n[i]; geo[9 + i] = Ae[i]; }
    xcur[xoff] = ac;
  }
}

AFF_DEV double k_region_dx(const KTask* tk, double e)
{
  if (!ldg(&tk->has_region)) return e;
  const double off = -e;
  const double rmax = ldg(&tk->region_max), rmin = ldg(&tk->region_min);
  if (off > rmax) return e + rmax;
  if (off < rmin) return e + rmin;
  return ldg(&tk->region_dx_scaling) * e;
}

// Task::computeDX towards xdes, from geo / xcur of the current state.
AFF_DEV void k_task_dx(const int* td, const KTask* tk, const double* geo, const double* xcur, const double* xdes, double* dx)
{
  const int kind = td[TD_KIND], xoff = td[TD_XOFF];
  if (kind == AFF_TASK_POLAR) {
    const double* a = geo; const double* e1 = geo + 3; const double* e2 = geo + 6;
    double ad[3], n[3], w[3], sp, cp, st, ct;
    aff_sincos(xdes[xoff], &sp, &cp);
    aff_sincos(xdes[xoff + 1], &st, &ct);
    ad[0] = sp * ct; ad[1] = sp * st; ad[2] = cp;
    k_cross(n, a, ad);
    const double s = sqrt(k_dot(n, n));
    const double cc = k_dot(a, ad);
    const double ang = aff_atan2(s, cc);
    if (s > 1.0e-12) { const double f = ang / s; for (int k = 0; k < 3; ++k) w[k] = n[k] * f; }
    else { w[0] = w[1] = w[2] = 0.0; }
    dx[0] = k_dot(e1, w);
    dx[1] = k_dot(e2, w);
    return;
  }
  const int dim = td[TD_DIM];
  for (int i = 0; i < dim; ++i) dx[i] = k_region_dx(tk, xdes[xoff + i] - xcur[xoff + i]);
}

// Rows of one task in column `lane` of J (Task::computeJ).  Returns a bit per
// row that is non-zero in this column.
template <class Ctx> AFF_DEV unsigned k_task_J(const Ctx& c, const KJointFrame& jf, int col, const int* td, const KTask* tk, const double* geo, double* Jrow0, int Jstride)
{
  const int kind = td[TD_KIND], eff = td[TD_EFF], ref = td[TD_REF], frame = td[TD_FRAME];
  unsigned nz = 0u;
  if (kind <= AFF_TASK_Z) {
    double te[3], re[3], tr[3], rr[3], tf[3], rf[3], cr[3], cl[3], out[3];
    k_jac_cols(c, jf, col, eff, geo, te, re);
    k_jac_cols(c, jf, col, ref, geo + 3, tr, rr);
    k_jac_cols(c, jf, col, frame, k_slot(c, frame >= 0 ? frame : 0), tf, rf);      // without a frame the point is not read
    k_cross(cr, geo + 6, rf);
    for (int k = 0; k < 3; ++k) cl[k] = (te[k] - tr[k]) + cr[k];
    if (frame >= 0) k_mulv(out, k_slot(c, frame) + 3, cl);
    else { out[0] = cl[0]; out[1] = cl[1]; out[2] = cl[2]; }
    if (kind == AFF_TASK_XYZ) {
      Jrow0[col] = out[0]; Jrow0[Jstride + col] = out[1]; Jrow0[2 * Jstride + col] = out[2];
      nz = (out[0] != 0.0 ? 1u : 0u) | (out[1] != 0.0 ? 2u : 0u) | (out[2] != 0.0 ? 4u : 0u);
    } else { const double v = k_sel3(out, kind - AFF_TASK_X); Jrow0[col] = v; nz = v != 0.0 ? 1u : 0u; }
  } else if (kind == AFF_TASK_POLAR) {
    double te[3], re[3], tr[3], rr[3], wI[3], w[3];
    k_jac_cols(c, jf, col, eff, geo + 9, te, re);
    k_jac_cols(c, jf, col, ref, geo + 9, tr, rr);
    for (int k = 0; k < 3; ++k) wI[k] = re[k] - rr[k];
    if (ref >= 0) k_mulv(w, k_slot(c, ref) + 3, wI);
    else { w[0] = wI[0]; w[1] = wI[1]; w[2] = wI[2]; }
    const double j0 = k_dot(geo + 3, w), j1 = k_dot(geo + 6, w);
    Jrow0[col] = j0; Jrow0[Jstride + col] = j1;
    nz = (j0 != 0.0 ? 1u : 0u) | (j1 != 0.0 ? 2u : 0u);
  } else if (kind == AFF_TASK_INCLINATION) {
    double te[3], re[3], tr[3], rr[3], wI[3];
    const double* n = geo + 6;
    double v = 0.0;
    if (n[0] != 0.0 || n[1] != 0.0 || n[2] != 0.0) {
      k_jac_cols(c, jf, col, eff, geo + 9, te, re);
      k_jac_cols(c, jf, col, ref, geo + 9, tr, rr);
      for (int k = 0; k < 3; ++k) wI[k] = re[k] - rr[k];
      v = k_dot(n, wI);
    }
    Jrow0[col] = v;
    nz = v != 0.0 ? 1u : 0u;
  } else {
    const int nj = td[TD_NJ];
    for (int i = 0; i < nj; ++i) { const bool on = ldg(&tk->jidx[i]) == col; Jrow0[i * Jstride + col] = on ? 1.0 : 0.0; if (on) nz |= 1u << i; }
  }
  return nz;
}

// -------------------------------------------------------- shape distances
AFF_DEV double k_seg_seg(const double* p1, const double* d1, const double* p2, const double* d2)
{
  const double EPS = 1.0e-18;
  double r[3];
  for (int k = 0; k < 3; ++k) r[k] = p1[k] - p2[k];
  const double a = k_dot(d1, d1), e = k_dot(d2, d2), f = k_dot(d2, r);
  double s, t;
  if (a <= EPS && e <= EPS) { s = 0.0; t = 0.0; }
  else if (a <= EPS) { s = 0.0; t = k_clamp01(f / e); }
  else {
    const double cc = k_dot(d1, r);
    if (e <= EPS) { t = 0.0; s = k_clamp01(-cc / a); }
    else {
      const double b = k_dot(d1, d2);
      const double denom = fma(a, e, -(b * b));
      s = (denom > 0.0) ? k_clamp01(fma(b, f, -(cc * e)) / denom) : 0.0;
      t = fma(b, s, f) / e;
      if (t < 0.0) { t = 0.0; s = k_clamp01(-cc / a); }
      else if (t > 1.0) { t = 1.0; s = k_clamp01((b - cc) / a); }
    }
  }
  double dv[3];
  for (int k = 0; k < 3; ++k) dv[k] = fma(s, d1[k], p1[k]) - fma(t, d2[k], p2[k]);
  return sqrt(k_dot(dv, dv));
}

AFF_DEV double k_seg_box_g(const double* p, const double* d, const double* h, double s)
{
  double g = 0.0;
  for (int i = 0; i < 3; ++i) {
    const double cv = fma(s, d[i], p[i]);
    const double e = (cv > h[i]) ? cv - h[i] : ((cv < -h[i]) ? cv + h[i] : 0.0);
    g = fma(e, d[i], g);
  }
  return g;
}
AFF_DEV double k_seg_box_f(const double* p, const double* d, const double* h, double s)
{
  double f = 0.0;
  for (int i = 0; i < 3; ++i) {
    const double cv = fma(s, d[i], p[i]);
    const double e = (cv > h[i]) ? cv - h[i] : ((cv < -h[i]) ? cv + h[i] : 0.0);
    f = fma(e, e, f);
  }
  return f;
}
AFF_DEV double k_seg_box(const double* p, const double* d, const double* h)
{
  double lo = 0.0, hi = 1.0;
  double glo = k_seg_box_g(p, d, h, 0.0), ghi = k_seg_box_g(p, d, h, 1.0);
  double s;
  if (glo >= 0.0) s = 0.0;
  else if (ghi <= 0.0) s = 1.0;
  else {
    for (int i = 0; i < 3; ++i) {
      if (d[i] == 0.0) continue;
      for (int sg = 0; sg < 2; ++sg) {
        const double b = ((sg ? h[i] : -h[i]) - p[i]) / d[i];
        if (!(b > lo && b < hi)) continue;
        const double gb = k_seg_box_g(p, d, h, b);
        if (gb <= 0.0) { lo = b; glo = gb; } else { hi = b; ghi = gb; }
      }
    }
    const double dg = ghi - glo;
    s = (dg > 0.0) ? fma(hi - lo, (-glo) / dg, lo) : lo;
  }
  return sqrt(k_seg_box_f(p, d, h, s));
}

AFF_DEV double k_seg_cyl_D(const double* p, const double* d, double hl, double R, double s)
{
  const double x = fma(s, d[0], p[0]), y = fma(s, d[1], p[1]), z = fma(s, d[2], p[2]);
  const double rho = sqrt(fma(y, y, x * x));
  const double er = (rho > R) ? rho - R : 0.0;
  const double az = fabs(z);
  const double ez = (az > hl) ? az - hl : 0.0;
  return sqrt(fma(ez, ez, er * er));
}
AFF_DEV double k_seg_cyl(const double* p, const double* d, double hl, double R)
{
  const double PHI = 0.6180339887498949;
  double best = k_seg_cyl_D(p, d, hl, R, 0.0);
  if (d[0] == 0.0 && d[1] == 0.0 && d[2] == 0.0) return best;
  const double D1 = k_seg_cyl_D(p, d, hl, R, 1.0);
  if (D1 < best) best = D1;
  double a = 0.0, b = 1.0;
  double x1 = b - PHI * (b - a), x2 = a + PHI * (b - a);
  double f1 = k_seg_cyl_D(p, d, hl, R, x1), f2 = k_seg_cyl_D(p, d, hl, R, x2);
  for (int it = 0; it < AFFK_GOLDEN_ITERS; ++it) {
    if (f1 <= f2) { b = x2; x2 = x1; f2 = f1; x1 = b - PHI * (b - a); f1 = k_seg_cyl_D(p, d, hl, R, x1); }
    else { a = x1; x1 = x2; f1 = f2; x2 = a + PHI * (b - a); f2 = k_seg_cyl_D(p, d, hl, R, x2); }
  }
  const double Dm = k_seg_cyl_D(p, d, hl, R, 0.5 * (a + b));
  if (Dm < best) best = Dm;
  return best;
}

AFF_DEV bool k_is_seg(int type) { return type == AFF_SHAPE_SSL || type == AFF_SHAPE_SPHERE; }

AFF_DEV void k_shape_segment(int type, const double* ext, const double* A, double* p, double* d, double* r)
{
  for (int k = 0; k < 3; ++k) p[k] = A[k];
  *r = ext[0];
  if (type == AFF_SHAPE_SSL) for (int k = 0; k < 3; ++k) d[k] = ext[2] * A[3 + 6 + k];
  else d[0] = d[1] = d[2] = 0.0;
}

// A1/A2: world shape frames (12 doubles); ext: 3 doubles each.
AFF_DEV double k_shape_dist(int t1, const double* e1, const double* A1, int t2, const double* e2, const double* A2)
{
  if (!k_is_seg(t1) && k_is_seg(t2)) {
    const int tt = t1; t1 = t2; t2 = tt;
    const double* te = e1; e1 = e2; e2 = te;
    const double* ta = A1; A1 = A2; A2 = ta;
  }
  if (!k_is_seg(t1)) return DBL_MAX;
  double p1[3], d1[3], r1;
  k_shape_segment(t1, e1, A1, p1, d1, &r1);
  if (k_is_seg(t2)) {
    double p2[3], d2[3], r2;
    k_shape_segment(t2, e2, A2, p2, d2, &r2);
    return (k_seg_seg(p1, d1, p2, d2) - r1) - r2;
  }
  double rel[3], pl[3], dl[3];
  for (int k = 0; k < 3; ++k) rel[k] = p1[k] - A2[k];
  k_mulv(pl, A2 + 3, rel);
  k_mulv(dl, A2 + 3, d1);
  if (t2 == AFF_SHAPE_BOX) {
    const double h[3] = {0.5 * e2[0], 0.5 * e2[1], 0.5 * e2[2]};
    return k_seg_box(pl, dl, h) - r1;
  }
  return k_seg_cyl(pl, dl, 0.5 * e2[2], e2[0]) - r1;
}

AFF_DEV void k_shape_aabb(int type, const double* ext, const double* A, double* mn, double* mx)
{
  if (type == AFF_SHAPE_SSL || type == AFF_SHAPE_SPHERE) {
    double p[3], d[3], r;
    k_shape_segment(type, ext, A, p, d, &r);
    for (int k = 0; k < 3; ++k) {
      const double q = p[k] + d[k];
      mn[k] = ((p[k] < q) ? p[k] : q) - r;
      mx[k] = ((p[k] > q) ? p[k] : q) + r;
    }
  } else if (type == AFF_SHAPE_BOX) {
    for (int k = 0; k < 3; ++k) {
      const double e = fma(fabs(A[3 + 6 + k]), 0.5 * ext[2], fma(fabs(A[3 + 3 + k]), 0.5 * ext[1], fabs(A[3 + k]) * (0.5 * ext[0])));
      mn[k] = A[k] - e; mx[k] = A[k] + e;
    }
  } else {
    for (int k = 0; k < 3; ++k) {
      const double az = A[3 + 6 + k];
      const double cc = 1.0 - az * az;
      const double e = fma(fabs(az), 0.5 * ext[2], ext[0] * sqrt(cc > 0.0 ? cc : 0.0));
      mn[k] = A[k] - e; mx[k] = A[k] + e;
    }
  }
}

AFF_DEV bool k_aabb_overlap(const double* A, const double* B, double thr)
{
  bool ov = true;
  for (int cI = 0; cI < 3; ++cI) if (A[cI] - thr > B[3 + cI] || B[cI] - thr > A[3 + cI]) ov = false;
  return ov;
}

// ------------------------------------------------------- collision model
struct KCollResult { double minDist; int minB1, minB2; double cost; int overflow; };

// toggleSlotPlus1: slot+1 if the shape's body IS a re-parentable object (the
// body a CollisionModelConstraint names), else 0.
template <class Ctx> AFF_DEV bool k_dshape_active(const Ctx& c, const KShape* sh, int toggleSlotPlus1)
{
  if (toggleSlotPlus1) {
    const int mode = OBJ_GET(c, objMode, toggleSlotPlus1 - 1);
    if (mode == 1) return true;
    if (mode == 2) return false;
  }
  return ldg(&sh->active0) != 0;
}

// Distance of a body pair found by the AABB test = min over active shape pairs.
// ia: dynamic body index; ib >= 0 dynamic body index, ib < 0 static body (-1-ib).
// (Rare path: only bodies whose boxes come within DistanceThreshold.)
template <class Ctx> AFF_DEV double k_pair_distance(const Ctx& c, int ia, int ib)
{
  const KPlan& P = *c.P;
  const KShapeBody* ba = P.dbody + ia;
  const int fa = ldg(&ba->first_shape), na = ldg(&ba->n_shapes), sa = ldg(&ba->toggle_slot);
  const double* shp = c.shpv();
  double dmin = DBL_MAX;
  for (int k1 = 0; k1 < na; ++k1) {
    const KShape* s1 = P.dshape + fa + k1;
    if (!k_dshape_active(c, s1, sa)) continue;
    double e1[3] = {ldg(&s1->ext[0]), ldg(&s1->ext[1]), ldg(&s1->ext[2])};
    const int t1 = ldg(&s1->type);
    if (ib >= 0) {
      const KShapeBody* bb = P.dbody + ib;
      const int fb = ldg(&bb->first_shape), nb = ldg(&bb->n_shapes), sb = ldg(&bb->toggle_slot);
      for (int k2 = 0; k2 < nb; ++k2) {
        const KShape* s2 = P.dshape + fb + k2;
        if (!k_dshape_active(c, s2, sb)) continue;
        double e2[3] = {ldg(&s2->ext[0]), ldg(&s2->ext[1]), ldg(&s2->ext[2])};
        const double dd = k_shape_dist(t1, e1, shp + 12 * (fa + k1), ldg(&s2->type), e2, shp + 12 * (fb + k2));
        if (dd < dmin) dmin = dd;
      }
    } else {
      const KShapeBody* bb = P.sbody + (-1 - ib);
      const int fb = ldg(&bb->first_shape), nb = ldg(&bb->n_shapes);
      const bool swapStatic = ldg(&bb->body) < ldg(&ba->body);
      for (int k2 = 0; k2 < nb; ++k2) {
        const KShape* s2 = P.sshape + fb + k2;
        if (!ldg(&s2->active0)) continue;
        double e2[3] = {ldg(&s2->ext[0]), ldg(&s2->ext[1]), ldg(&s2->ext[2])};
        double A2[12];
        for (int e = 0; e < 12; ++e) A2[e] = ldg(&P.sshapeA[12 * (fb + k2) + e]);
        // argument order = ascending scene body id, as the CPU checker enumerates pairs
        const double dd = swapStatic ? k_shape_dist(ldg(&s2->type), e2, A2, t1, e1, shp + 12 * (fa + k1))
                                     : k_shape_dist(t1, e1, shp + 12 * (fa + k1), ldg(&s2->type), e2, A2);
        if (dd < dmin) dmin = dd;
      }
    }
  }
  return dmin;
}

// meta word of a precompiled shape pair (plan_build.cpp: packPairMeta)
#define KP_SLOT_A(m) ((m) & 3)
#define KP_SLOT_B(m) (((m) >> 2) & 3)
#define KP_SWAP(m) (((m) >> 4) & 1)
#define KP_TREE_A(m) ((int)(((m) >> 9) & 15) - 1)
#define KP_TREE_B(m) ((int)(((m) >> 13) & 15) - 1)
#define KP_EXPL_A(m) (((m) >> 17) & 1)
#define KP_EXPL_B(m) (((m) >> 18) & 1)
#define KP_ACT_A(m) (((m) >> 19) & 1)
#define KP_ACT_B(m) (((m) >> 20) & 1)
#define KP_BODY_A(m) (((m) >> 21) & 31)
#define KP_BODY_B(m) (((m) >> 26) & 31)    /* 31 = static body */
#define KP_RULE(m) (((unsigned)(m)) >> 31)  /* liveness must be decided per step */

// Is a precompiled pair part of this step's collision model?  Tree x tree and
// tree x explicit pairs of fixed membership always are (no RULE bit).  Pairs
// involving a re-parentable object depend on its shape flags and current tree,
// and a tree body pairs with a non-explicit body outside all trees only while
// their AABBs (inflated by the threshold) overlap.
template <class Ctx> AFF_DEV bool k_pair_live(const Ctx& c, int meta, const double* baabb, double thr)
{
  if (!KP_RULE(meta)) return true;
  const int sa = KP_SLOT_A(meta), sb = KP_SLOT_B(meta);
  bool actA = KP_ACT_A(meta), actB = KP_ACT_B(meta);
  int ta = KP_TREE_A(meta), tb = KP_TREE_B(meta);
  if (sa) { const int mode = OBJ_GET(c, objMode, sa - 1); if (mode == 1) actA = true; else if (mode == 2) actA = false; ta = OBJ_GET(c, objTree, sa - 1); }
  if (sb) { const int mode = OBJ_GET(c, objMode, sb - 1); if (mode == 1) actB = true; else if (mode == 2) actB = false; tb = OBJ_GET(c, objTree, sb - 1); }
  if (!actA || !actB) return false;
  if (ta < 0 && tb < 0) return false;
  if (ta >= 0 && tb >= 0) return ta != tb;
  const bool otherExplicit = (ta >= 0) ? KP_EXPL_B(meta) : KP_EXPL_A(meta);
  if (otherExplicit) return true;
  const int ba = KP_BODY_A(meta), bb = KP_BODY_B(meta);
  if (bb == 31) return false;   // static, not explicit: handled by the AABB path
  return k_aabb_overlap(baabb + 6 * ba, baabb + 6 * bb, thr);
}

// Fold one shape-pair distance into the lane's running minimum; pairs closer than the
// threshold (rare) are appended to a scratch list for the ordered cost sum.
struct KPairAcc { double best; int bestKey; int nCostly; };

AFF_DEV void k_pair_account(KPairAcc& acc, int lane, bool on, double d, int key, double thr, int* ckey, double* cdist)
{
  if (on && (d < acc.best || (d == acc.best && key < acc.bestKey))) { acc.best = d; acc.bestKey = key; }
  const int costly = on && thr > 0.0 && d < thr;
  const unsigned m = w_ballot(costly);
  if (m) {
    if (costly) {
      const int pos = acc.nCostly + w_popc(m & ((1u << lane) - 1u));
      if (pos < 32) { ckey[pos] = key; cdist[pos] = d; }
    }
    acc.nCostly += w_popc(m);
  }
}

// ControllerBase::computeCollisionModel: broad phase + narrow phase
// (see the rule list above collmdl_compute() in oracle/oracle.c).
// `need`: the caller only looks at the minimum if it is below this value (the running minimum of
// the rollout or the 1 mm collision limit), so the warp-wide arg-min is skipped otherwise.
AFF_DEV KCollResult k_collision(const KCtx& c, double need)
{
  const KPlan& P = *c.P;
  const int lane = c.lane;
  const double thr = P.bp_threshold;
  double* shp = c.sm + AFFK_O(P, shp);
  double* baabb = c.sm + AFFK_O(P, baabb);
  double* taabb = c.sm + AFFK_O(P, taabb);
  int* live = (int*)(c.sm + AFFK_O(P, live));             // AABB-found pairs: [2*i] = dyn body a, [2*i+1] = other
  int* btree = live + 2 * AFFK_MAX_NEAR;           // tree of each dynamic body, -1 none, -2 inactive
  int* ckey = btree + AFFK_MAX_DBODY;              // scratch: pairs closer than the threshold
  double* cdist = (double*)(ckey + 32);
  // 1. world records of the dynamic shapes: org (0-2), SSL direction d = length * z-axis (3-5),
  //    z-axis (9-11); boxes and cylinders get the full frame.
  for (int s = lane; s < P.nDynShapes; s += 32) {
    const KShape* sh = P.dshape + s;
    const double* An = k_slot(c, ldg(&sh->node));
    const int type = ldg(&sh->type);
    double* W = shp + 12 * s;
    if (type == AFF_SHAPE_SSL || type == AFF_SHAPE_SPHERE) {
      for (int e = 0; e < 3; ++e) W[e] = k_compose_elem(sh->A_CB, An, e);
      if (type == AFF_SHAPE_SSL) {
        const double len = ldg(&sh->ext[2]);
        for (int e = 0; e < 3; ++e) { const double ax = k_compose_elem(sh->A_CB, An, 9 + e); W[9 + e] = ax; W[3 + e] = len * ax; }
      } else { W[3] = 0.0; W[4] = 0.0; W[5] = 0.0; }
      W[6] = ldg(&sh->ext[0]); W[7] = ldg(&sh->bound);
    } else {
      for (int e = 0; e < 12; ++e) W[e] = k_compose_elem(sh->A_CB, An, e);
    }
  }
  w_sync();
  // 2. per dynamic body: active flag, tree, AABB
  if (lane < P.nDynBodies) {
    const KShapeBody* b = P.dbody + lane;
    const int f = ldg(&b->first_shape), n = ldg(&b->n_shapes), os = ldg(&b->obj_slot), ts = ldg(&b->toggle_slot);
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    int act = 0;
    for (int k = 0; k < n; ++k) {
      const KShape* sh = P.dshape + f + k;
      if (!k_dshape_active(c, sh, ts)) continue;
      act = 1;
      const int type = ldg(&sh->type);
      const double* W = shp + 12 * (f + k);
      if (type == AFF_SHAPE_SSL || type == AFF_SHAPE_SPHERE) {
        const double r = ldg(&sh->ext[0]);
        for (int q = 0; q < 3; ++q) {
          const double p0 = W[q], p1 = p0 + W[3 + q];
          const double lo = ((p0 < p1) ? p0 : p1) - r, hi = ((p0 > p1) ? p0 : p1) + r;
          if (lo < mn[q]) mn[q] = lo;
          if (hi > mx[q]) mx[q] = hi;
        }
      } else {
        double ext[3] = {ldg(&sh->ext[0]), ldg(&sh->ext[1]), ldg(&sh->ext[2])};
        double smn[3], smx[3];
        k_shape_aabb(type, ext, W, smn, smx);
        for (int q = 0; q < 3; ++q) { if (smn[q] < mn[q]) mn[q] = smn[q]; if (smx[q] > mx[q]) mx[q] = smx[q]; }
      }
    }
    for (int q = 0; q < 3; ++q) { baabb[6 * lane + q] = mn[q]; baabb[6 * lane + 3 + q] = mx[q]; }
    btree[lane] = act ? (os ? OBJ_GET(c, objTree, os - 1) : ldg(&b->tree)) : -2;
  }
  w_sync();
  // 3. tree AABBs: one lane per (tree, bound); min / max over the tree's bodies
  for (int i = lane; i < 6 * P.nTrees; i += 32) {
    const int t = i / 6, k = i % 6;
    const bool isMin = k < 3;
    double v = isMin ? DBL_MAX : -DBL_MAX;
    for (int b = 0; b < P.nDynBodies; ++b) {
      const double x = baabb[6 * b + k];
      const bool take = (btree[b] == t) && (isMin ? x < v : x > v);
      v = take ? x : v;
    }
    taabb[i] = v;
  }
  w_sync();
  // 4. static bodies that are not explicit: two-level AABB test (tree box, then tree bodies)
  int nLive = 0;
  for (int base = 0; base < P.nStatNear; base += 32) {
    const int i = base + lane;
    int near = 0;
    if (i < P.nStatNear) {          // the plan lists the bodies this test applies to first
      double bb[6];
      for (int q = 0; q < 6; ++q) bb[q] = ldg(&P.sbodyAABB[6 * i + q]);
      for (int t = 0; t < P.nTrees; ++t) if (k_aabb_overlap(taabb + 6 * t, bb, thr)) near = 1;
    }
    unsigned nm = w_ballot(near);
    while (nm) {   // rare: a static body is close to a tree; test it against each tree body
      const int src = w_ffs(nm) - 1;
      nm &= nm - 1u;
      const int sidx = base + src;
      int isLive = 0;
      if (lane < P.nDynBodies && btree[lane] >= 0) {
        double bb[6];
        for (int q = 0; q < 6; ++q) bb[q] = ldg(&P.sbodyAABB[6 * sidx + q]);
        isLive = k_aabb_overlap(baabb + 6 * lane, bb, thr);
      }
      const unsigned m = w_ballot(isLive);
      if (isLive) {
        const int pos = nLive + w_popc(m & ((1u << lane) - 1u));
        if (pos < AFFK_MAX_NEAR) { live[2 * pos] = lane; live[2 * pos + 1] = -1 - sidx; }
      }
      nLive += w_popc(m);
    }
  }
  const int nearOverflow = nLive > AFFK_MAX_NEAR;
  if (nLive > AFFK_MAX_NEAR) nLive = AFFK_MAX_NEAR;
  // 5. narrow phase over the precompiled, type-sorted shape-pair lists
  KPairAcc acc;
  acc.best = DBL_MAX; acc.bestKey = 0x7fffffff; acc.nCostly = 0;
  // A pair whose bounding spheres are farther apart than `need` cannot become the minimum the caller
  // looks at, cannot collide and cannot cost (thr <= 1 mm <= need): its exact distance is skipped.
  const double cullMargin = 1.0e-9;
  const double cullNeed = (need > thr) ? need : thr;     // pairs closer than the threshold carry a cost term
  // 5a. segment - segment (capsules, spheres)
  for (int base = 0; base < P.nSS; base += 32) {
    const int i = base + lane;
    double d = DBL_MAX; int key = 0x7fffffff; bool on = false;
    if (i < P.nSS) {
      const KPair pr = ldg_pair(&P.pairsSS[i]);
      key = pr.key;
      if (k_pair_live(c, pr.meta, baabb, thr)) {
        const double* A = shp + 12 * pr.a;
        const double rA = A[6];
        double pB[3], dB[3], rB, bB;
        if (pr.b >= 0) { const double* B = shp + 12 * pr.b; for (int k = 0; k < 3; ++k) { pB[k] = B[k]; dB[k] = B[3 + k]; } rB = B[6]; bB = B[7]; }
        else { const double* B = P.sseg + 8 * (-1 - pr.b); for (int k = 0; k < 3; ++k) { pB[k] = ldg(&B[k]); dB[k] = ldg(&B[3 + k]); } rB = ldg(&B[6]); bB = ldg(&B[7]); }
        double cd2 = 0.0;
        for (int k = 0; k < 3; ++k) { const double dc = (A[k] + 0.5 * A[3 + k]) - (pB[k] + 0.5 * dB[k]); cd2 += dc * dc; }
        const double lim = cullNeed + (A[7] + bB) + cullMargin;
        if (!(cd2 >= lim * lim)) {
          on = true;
          d = KP_SWAP(pr.meta) ? (k_seg_seg(pB, dB, A, A + 3) - rB) - rA : (k_seg_seg(A, A + 3, pB, dB) - rA) - rB;
        }
      }
    }
    k_pair_account(acc, lane, on, d, key, thr, ckey, cdist);
  }
  // 5b. segment - box and segment - cylinder (the segment is always the first argument)
  for (int base = 0; base < P.nSB; base += 32) {
    const int i = base + lane;
    double d = DBL_MAX; int key = 0x7fffffff; bool on = false;
    if (i < P.nSB) {
      const KPair pr = ldg_pair(&P.pairsSB[i]);
      key = pr.key;
      if (k_pair_live(c, pr.meta, baabb, thr)) {
        // a = segment shape (dynamic >= 0 / static < 0), b = box or cylinder shape
        double pS[3], dS[3], rS, bS, cB[3], bB;
        if (pr.a >= 0) { const double* A = shp + 12 * pr.a; for (int k = 0; k < 3; ++k) { pS[k] = A[k]; dS[k] = A[3 + k]; } rS = A[6]; bS = A[7]; }
        else { const double* A = P.sseg + 8 * (-1 - pr.a); for (int k = 0; k < 3; ++k) { pS[k] = ldg(&A[k]); dS[k] = ldg(&A[3 + k]); } rS = ldg(&A[6]); bS = ldg(&A[7]); }
        if (pr.b >= 0) { const double* B = shp + 12 * pr.b; for (int k = 0; k < 3; ++k) cB[k] = B[k]; bB = ldg(&P.dshape[pr.b].bound); }
        else { const double* B = P.sseg + 8 * (-1 - pr.b); for (int k = 0; k < 3; ++k) cB[k] = ldg(&B[k]); bB = ldg(&B[7]); }
        double cd2 = 0.0;
        for (int k = 0; k < 3; ++k) { const double dc = (pS[k] + 0.5 * dS[k]) - cB[k]; cd2 += dc * dc; }
        const double lim = cullNeed + (bS + bB) + cullMargin;
        if (!(cd2 >= lim * lim)) {
          on = true;
          // segment into the box / cylinder frame, one row of the frame at a time (few live registers)
          const bool dynB = pr.b >= 0;
          const double* Bs = shp + 12 * (dynB ? pr.b : 0);
          const double* Bg = P.sshapeA + 12 * (dynB ? 0 : -1 - pr.b);
          const KShape* sb = dynB ? P.dshape + pr.b : P.sshape + (-1 - pr.b);
          double rel[3], pl[3], dl[3];
          for (int k = 0; k < 3; ++k) rel[k] = pS[k] - (dynB ? Bs[k] : ldg(&Bg[k]));
          for (int i = 0; i < 3; ++i) {
            const double r0 = dynB ? Bs[3 + 3 * i] : ldg(&Bg[3 + 3 * i]), r1 = dynB ? Bs[4 + 3 * i] : ldg(&Bg[4 + 3 * i]),
                         r2 = dynB ? Bs[5 + 3 * i] : ldg(&Bg[5 + 3 * i]);
            pl[i] = fma(r2, rel[2], fma(r1, rel[1], r0 * rel[0]));
            dl[i] = fma(r2, dS[2], fma(r1, dS[1], r0 * dS[0]));
          }
          const int typeB = ldg(&sb->type);
          double ext[3];
          for (int k = 0; k < 3; ++k) ext[k] = ldg(&sb->ext[k]);
          if (typeB == AFF_SHAPE_BOX) { const double h[3] = {0.5 * ext[0], 0.5 * ext[1], 0.5 * ext[2]}; d = k_seg_box(pl, dl, h) - rS; }
          else d = k_seg_cyl(pl, dl, 0.5 * ext[2], ext[0]) - rS;
        }
      }
    }
    k_pair_account(acc, lane, on, d, key, thr, ckey, cdist);
  }
  // 5c. pairs found by the AABB test (rare)
  for (int base = 0; base < nLive; base += 32) {
    const int i = base + lane;
    double d = DBL_MAX; int key = 0x7fffffff;
    if (i < nLive) {
      const int ia = live[2 * i], ib = live[2 * i + 1];
      d = k_pair_distance(c, ia, ib);
      const int bodyA = ldg(&P.dbody[ia].body);
      const int bodyB = (ib >= 0) ? ldg(&P.dbody[ib].body) : ldg(&P.sbody[-1 - ib].body);
      const int b1 = bodyA < bodyB ? bodyA : bodyB, b2 = bodyA < bodyB ? bodyB : bodyA;
      key = (b1 << 15) | b2;
    }
    k_pair_account(acc, lane, i < nLive, d, key, thr, ckey, cdist);
  }
  KCollResult res;
  double best = acc.best; int bestKey = acc.bestKey;
  if (w_ballot(best < need)) {
    w_min_keyed(best, bestKey);
    res.minDist = best;
    if (!(best < DBL_MAX)) { res.minB1 = -1; res.minB2 = -1; }
    else { res.minB1 = bestKey >> 15; res.minB2 = bestKey & 0x7fff; }
  } else { res.minDist = DBL_MAX; res.minB1 = -1; res.minB2 = -1; }
  // cost: one term per BODY pair (its closest shape pair), added in ascending (b1, b2) order.
  // Only pairs closer than the threshold contribute, so this loop almost never runs.
  res.cost = 0.0;
  res.overflow = (nearOverflow || acc.nCostly > 32) ? 1 : 0;
  if (acc.nCostly > 0) {
    w_sync();
    const int n = acc.nCostly < 32 ? acc.nCostly : 32;
    int myKey = (lane < n) ? ckey[lane] : 0x7fffffff;
    const double myD = (lane < n) ? cdist[lane] : DBL_MAX;
    for (;;) {
      int kmin = myKey;
      for (int off = 16; off >= 1; off >>= 1) { const int o = w_shfl_xor(kmin, off); kmin = o < kmin ? o : kmin; }
      if (kmin == 0x7fffffff) break;
      const double d = w_min(myKey == kmin ? myD : DBL_MAX);
      res.cost = res.cost + ((d < 0.0) ? 1.0 - (2.0 * d) / thr : ((d - thr) / thr) * ((d - thr) / thr));
      if (myKey == kmin) myKey = 0x7fffffff;
    }
  }
  return res;
}

// dH of one step (:589-625): joint-limit gradient + elbow / wrist null-space terms, scaled and speed-limited;
// lane = IK joint.  Reads the frames, the chains and q of the state before the step.
template <class Ctx> AFF_DEV double k_null_space_dH(const Ctx& c, const KJointFrame& jf, const double* q, double alphaEff, double phaseScale)
{
  const KPlan& P = *c.P;
  const int lane = c.lane, nJ = P.nJ;
  const double dt = P.opts.dt;
  const int jl = (lane < nJ) ? lane : 0;
  const double jq0 = ldg_here(&P.jq0[jl]), jqmin = ldg_here(&P.jq_min[jl]), jqmax = ldg_here(&P.jq_max[jl]);
  const double jwjl = ldg_here(&P.jwjl[jl]), jspeed = ldg_here(&P.jspeed[jl]);
  double dHv = 0.0;
  if (lane < nJ) {
    const double dqc = q[lane] - jq0;
    const double r = (dqc > 0.0) ? jqmax - jq0 : jq0 - jqmin;
    dHv = ((jwjl * dqc) / (r * r)) * alphaEff;
  }
  {
    double ex = 0.0, tmp = 0.0;
    if (P.ns_base >= 0) {   // -1: base_footprint absent, no extra null space
      const int base = P.ns_base;
      const double* Ab = k_slot(c, base);
      // elbows, then hands (boundary 0.15 / 0.0)
      for (int pass = 0; pass < 2; ++pass) {
        const double boundary = pass ? 0.0 : 0.15;
        for (int side = 0; side < 2; ++side) {
          const int el = pass ? P.ns_hand[side] : P.ns_forearm[side];
          const int sh = P.ns_upperarm[side];
          if (el < 0 || sh < 0) continue;
          const double* Ae = k_slot(c, el); const double* As = k_slot(c, sh);
          double dS[3], dE[3];
          for (int k = 0; k < 3; ++k) { dS[k] = As[k] - Ab[k]; dE[k] = Ae[k] - Ab[k]; }
          const double yS = fma(Ab[3 + 3 + 2], dS[2], fma(Ab[3 + 3 + 1], dS[1], Ab[3 + 3] * dS[0]));
          const double yE = fma(Ab[3 + 3 + 2], dE[2], fma(Ab[3 + 3 + 1], dE[1], Ab[3 + 3] * dE[0]));
          double f;
          if (side == 0) f = k_clip((yE - yS) - boundary, -0.5, 0.0) * 5.0;
          else f = -k_clip((-yE + yS) - boundary, -0.5, 0.0) * 5.0;
          // f == +-0 (the usual case: elbow / hand clear of its boundary): the term is +-0 and leaves the
          // sum, which is never -0, as it is; f is the same on every lane
          if (f != 0.0) {
            double te[3], re[3], tr[3], rr[3];
            k_jac_cols(c, jf, lane, el, Ae, te, re);
            k_jac_cols(c, jf, lane, base, Ab, tr, rr);
            tmp = tmp + (te[1] - tr[1]) * f;
          }
        }
      }
      ex = ex + tmp;
      tmp = 0.0;
      for (int side = 0; side < 2; ++side) {
        const int wr = P.ns_hand[side];
        if (wr < 0) continue;
        const double* Aw = k_slot(c, wr);
        double dW[3];
        for (int k = 0; k < 3; ++k) dW[k] = Aw[k] - Ab[k];
        const double x = fma(Ab[3 + 2], dW[2], fma(Ab[3 + 1], dW[1], Ab[3] * dW[0]));
        const double z = fma(Ab[3 + 6 + 2], dW[2], fma(Ab[3 + 6 + 1], dW[1], Ab[3 + 6] * dW[0]));
        const double dBound = (z > 1.0) ? 0.4 * (z - 1.0) : 0.0;
        const double f = k_clip(x - (0.25 + dBound), -0.5, 0.0) * 5.0;
        if (f != 0.0) {
          double te[3], re[3], tr[3], rr[3];
          k_jac_cols(c, jf, lane, wr, Aw, te, re);
          k_jac_cols(c, jf, lane, base, Ab, tr, rr);
          tmp = tmp + (te[0] - tr[0]) * f;
        }
      }
      ex = (ex + tmp) * alphaEff;
    }
    // RcsGraph_checkJointSpeeds(dH_extra, dt, IK)
    double sc = 1.0;
    if (lane < nJ && jspeed < AFFK_NO_LIMIT) { const double a = fabs(ex), mx = jspeed * dt; if (a > mx) sc = mx / a; }
    if (w_ballot(sc < 1.0)) { const double scale = w_min(sc); ex = ex * (0.99999 * scale); }
    dHv = dHv + ex * phaseScale;
  }
  return dHv;
}

AFF_DEV bool k_geo_helper_kind(int kind) { return kind == AFF_TASK_POLAR || kind == AFF_TASK_INCLINATION; }

// ----------------------------------------------------------- the rollout
// Control block of the block-per-candidate kernel (cta.cuh), in shared memory: warp 0 walks the step
// loop, the other warps evaluate the collision model of the steps it has finished, a few steps behind.
#define AFFC_WORKERS 2
#define AFFC_RING 4
#define AFFC_TRING 8       // steps the trajectory warp may run ahead
// warp roles: 0 step loop | 1 .. AFFC_WORKERS collision | AFFC_WORKERS + 1 trajectories | 5 null space | 6 late FK rounds | 7 geometry | 4 idle.
// Warps w and w + 4 share a scheduler: the step loop keeps its own to itself (a helper spinning on a flag
// beside it costs it 3 % - measured with the geometry warp as warp 4).
#define AFFC_FK_WARP 6
#define AFFC_GEO_WARP 7
#define AFFC_NS_WARP 5
#define AFFC_WARPS 8
// What the step loop needs from the time-driven part of a step (trajectories, activation points, blending):
// produced by the trajectory warp, which depends on the rollout only where a task switches on (its
// trajectory then starts from the current task value).
struct KCtaStep
{
  double xdes[32];
  double alphaEff, phaseScale, qFilt;
  unsigned activeMask; int pad;
};
struct KCtaShared
{
  int trajSeq;               // steps whose KCtaStep record is posted
  int geoSeq;                // 1 + number of steps whose task values (xcur) are final in warp 0's slab
  int consSeq;               // steps whose graph constraints are applied (chains, object state) and whose start state is final
  int nsSeq;                 // steps whose dH is posted by the null-space warp
  int trkSeq;                // steps whose tracking error is posted by the null-space warp
  int geoHelpSeq;            // steps whose orientation-task values are written by the geometry warp
  double dH[32];
  double trkDx[32];          // scratch of the tracking error
  KCtaStep ring[AFFC_TRING];
  int fkSeq;                 // number of steps whose frames of the rounds [0, nRoundsCrit) are final in warp 0's node array
  int fk2Seq;                // ... and of the remaining rounds (FK warp): all frames final
  int stop;                  // warp 0 has left the step loop
  int copied[AFFC_WORKERS];  // per worker: steps whose frames it has taken over
  int done[AFFC_WORKERS];    // per worker: steps whose result is posted
  int objState[4];           // objMode0, objMode1, objTree0, objTree1 of the posted step
  double need;               // running minimum distance (a stale, larger value only costs work)
  KCollResult res[AFFC_RING];
  // bookkeeping of steps whose collision result has not arrived yet
  int pendIk[AFFC_RING], pendId0[AFFC_RING], pendEk[AFFC_RING];
  double pendEv[AFFC_RING];
  int pendChk[AFFC_RING], pendChkId0[AFFC_RING];     // speed / joint-limit check of the state after the step (null-space warp)
  double pendJl[AFFC_RING];                            // joint-limit cost summed over the steps up to it
};

// CTA = 0: the warp does everything (one rollout per warp).  CTA = 1: warp 0 of a block-per-candidate
// launch; the collision model of step k is evaluated by a worker warp and folded into the verdict
// `lag` steps later.  The rollout itself never reads a collision result, so both orders give the same record.
template <int CTA> AFF_DEV void k_rollout_t(const KPlan& P, int cand, double* sm, KCtaShared* X)
{
  KCtx c;
  c.P = &P; c.C = P.cands + cand; c.sm = sm; c.lane = w_lane();
  c.nodes = sm + P.o_node;
  c.chain = (unsigned*)(sm + AFFK_O(P, chain));
  c.tdesc = (int*)(sm + AFFK_O(P, tdesc));
  const KCand& C = *c.C;
  const int lane = c.lane, nJ = P.nJ;
  const double dt = P.opts.dt;
  const int nTasks = ldg(&C.n_tasks), xdim = ldg(&C.xdim);
  const int taskBase = ldg(&C.first_task);
#define tasks (P.tasks + taskBase)
  double* q = sm + AFFK_O(P, q); double* qdot = sm + AFFK_O(P, qdot); double* vbuf = sm + AFFK_O(P, dH); double* wbuf = sm + AFFK_O(P, dq);
  double* J = sm + AFFK_O(P, J); double* L = sm + AFFK_O(P, L); double* rhs = sm + AFFK_O(P, rhs); double* dxv = sm + AFFK_O(P, dx);
  double* st = sm + AFFK_O(P, st); double* xdes = sm + AFFK_O(P, xdes); double* xcur = sm + AFFK_O(P, xcur); double* geo = sm + AFFK_O(P, geo);
  unsigned* rmask = (unsigned*)(sm + AFFK_O(P, rmask));
  const int JS = AFFK_JS(P), LS = AFFK_LS(P);
  const int nSlots = P.nDyn + P.nStatPar;

  // ---- per-lane constants
  // per-joint constants: re-read where they are used (ldg_here), not held in registers
  const int jl = (lane < nJ) ? lane : 0;
#define jq0 ldg_here(&P.jq0[jl])
#define jqmin ldg_here(&P.jq_min[jl])
#define jqmax ldg_here(&P.jq_max[jl])
#define jspeed ldg_here(&P.jspeed[jl])
#define jacc ldg_here(&P.jacc[jl])
#define jwjl ldg_here(&P.jwjl[jl])
#define jw ldg_here(&P.jwmetric[jl])
  int jnode = 0, jtype = 0;
  if (lane < nJ) {
    jnode = ldg(&P.j_node[lane]); jtype = ldg(&P.j_type[lane]);
    q[lane] = ldg(&P.jq_init[lane]); qdot[lane] = 0.0; wbuf[lane] = jw;
  }
  // task descriptors -> shared memory
  if (lane < nTasks) {
    const KTask* tk = tasks + lane;
    int* td = c.tdesc + 8 * lane;
    const int kind = ldg(&tk->kind);
    td[TD_KIND] = kind; td[TD_EFF] = ldg(&tk->eff); td[TD_REF] = ldg(&tk->ref); td[TD_FRAME] = ldg(&tk->frame);
    td[TD_AXIS] = ldg(&tk->axis); td[TD_XOFF] = ldg(&tk->xoff); td[TD_NJ] = ldg(&tk->n_joints);
    td[TD_DIM] = (kind == AFF_TASK_XYZ) ? 3 : ((kind == AFF_TASK_POLAR) ? 2 : ((kind == AFF_TASK_JOINTS) ? ldg(&tk->n_joints) : 1));
  }
  // task of this lane's dimension
  int myTask = -1;
  if (lane < xdim) for (int t = 0; t < nTasks; ++t) if (lane >= ldg(&tasks[t].xoff)) myTask = t;
  int viaHead = 0;      // the ends are re-read from the candidate record where needed
  // heads and ends are absolute indices into P.vias / P.acts: the table bases come from the constant bank
  if (lane < xdim) viaHead = ldg(&C.first_via) + ldg(&C.via_off[lane]);
#define viaEnd (ldg(&C.first_via) + ldg(&C.via_off[(lane < xdim ? lane : 0) + 1]))
#define vias P.vias
  int actHead = 0;      // last switch time / horizon / state are those of acts[actHead - 1]
  if (lane < nTasks) actHead = ldg(&C.first_act) + ldg(&C.act_off[lane]);
#define actBeginOf(t) (ldg(&C.first_act) + ldg(&C.act_off[t]))
#define actEndOf(t) (ldg(&C.first_act) + ldg(&C.act_off[(t) + 1]))
#define acts P.acts
  const int gconBase = ldg(&C.first_gcon);
#define gcons (P.gcons + gconBase)
  const int nGcon = ldg(&C.n_gcon);
  unsigned gfired = 0u;

  // ---- slots: static frames the loop refers to, chains
  for (int i = lane; i < 12 * P.nStatPar; i += 32) {
    const int k = i / 12, e = i % 12;
    const int ref = ldg(&P.statpar_ref[k]);
    const double* src = (ref == AFFK_WORLD) ? P.statA + 12 * P.nStatic : P.statA + 12 * AFFK_STATIC_IDX(ref);
    c.nodes[12 * (P.nDyn + k) + e] = ldg(&src[e]);
  }
  for (int i = lane; i < nSlots; i += 32) c.chain[i] = (i < P.nDyn) ? ldg(&P.dyn_chain[i]) : 0u;
  w_sync();
  // ---- objects
  c.objHasRel = 0u;
  c.objParentSlot0 = c.objParentSlot1 = 0; c.objParentDynamic0 = c.objParentDynamic1 = 0; c.objChain0 = c.objChain1 = 0u;
  c.objTree0 = c.objTree1 = -1; c.objMode0 = c.objMode1 = 0;
  for (int s = 0; s < P.nObj; ++s) {
    const int ps = ldg(&P.obj_parent_slot[s]);
    OBJ_SET(c, objParentSlot, s, ps);
    OBJ_SET(c, objParentDynamic, s, ps < P.nDyn);
    OBJ_SET(c, objChain, s, c.chain[ps]);             // held in a hand at start: the hand's chain
    OBJ_SET(c, objTree, s, ldg(&P.obj_parent_tree[s]));
  }
  w_sync();
  for (int i = lane; i < P.nDyn; i += 32) { const int uo = ldg(&P.dyn_under[i]); if (uo) c.chain[i] = ldg(&P.dyn_chain[i]) | OBJ_GET(c, objChain, uo - 1); }
  if (lane < 6) {
    for (int s = 0; s < P.nObj; ++s)
      sm[AFFK_O(P, objq6) + 6 * s + lane] = (ldg(&C.pose_slot) == s) ? ldg(&C.pose_q6[lane]) : ldg(&P.obj_q6[6 * s + lane]);
  }
  if (lane < xdim) { st[3 * lane] = 0.0; st[3 * lane + 1] = 0.0; st[3 * lane + 2] = 0.0; xdes[lane] = 0.0; }
  w_sync();

  // ---- initial state
  k_fk(c, true);
  const int nSteps = ldg(&C.n_steps);
  const int qrows = P.qlog_rows;
  double* qlog = P.qlog ? P.qlog + (size_t)cand * qrows * nJ : nullptr;
  if (qlog && lane < nJ && qrows > 0) qlog[lane] = q[lane];

  // result fields (warp-uniform scalars; written once at the end)
  int r_code = 0, r_first_code = 0, r_first_fail = -1, r_fail_step = -1, r_id0 = -1, r_id1 = -1, r_b1 = -1, r_b2 = -1;
  double r_minDist = DBL_MAX;

  KCollResult cm = k_collision(c, DBL_MAX);
  r_minDist = cm.minDist; r_b1 = cm.minB1; r_b2 = cm.minB2;
  unsigned r_flags = cm.overflow ? AFF_FLAG_LIST_OVERFLOW : 0u;

  // check(): checkLimits(jl, coll, speed, 0, 0, 0.01)  (src/TrajectoryPredictor.cpp:426-441)
  {
    const int viol = (lane < nJ) && (q[lane] < jqmin || q[lane] > jqmax);
    bool ok0 = w_ballot(viol) == 0u;
    if (cm.minB1 >= 0 && cm.minDist < 0.01) ok0 = false;
    if (!ok0) {
      if (lane == 0) {
        aff_result res;
        res.idx = cand; res.success = 0; res.code = AFF_CODE_INITIAL_STATE; res.first_code = AFF_CODE_INITIAL_STATE;
        res.first_fail_step = 0; res.fail_step = 0; res.fail_id0 = -1; res.fail_id1 = -1;
        res.min_dist_b1 = r_b1; res.min_dist_b2 = r_b2; res.n_steps = 0; res.reserved = r_flags; res.jmask_bits = 0ull;
        res.min_dist = r_minDist; res.jl_cost = 0.0; res.coll_cost = 0.0; res.track_err = 0.0;
        P.results[cand] = res;
      }
      if (CTA) w_post(&X->stop, 1);
      return;
    }
  }

  // geometry of the initial state (for the first re-initialisation of the trajectories)
  if (lane < nTasks) k_task_geo(c, c.tdesc + 8 * lane, tasks + lane, geo + AFFK_GEO * lane, xcur);
  w_sync();
  if (CTA) w_post(&X->geoSeq, 1);

  unsigned activeMask = 0u, prevMask = 0u, jmaskBits = 0u;
  unsigned emask = ~(P.rbj_static_chain);
  emask &= ~(c.objChain0 | c.objChain1);
  double tnow = 0.0, qFilt = 0.0, err = 0.0, jlSum = 0.0, collSum = 0.0, endTime = ldg(&C.end_abs);
  const double endAbs = ldg(&C.end_abs);
  const double motionDuration = endAbs - ldg(&C.start_abs);
  const double alpha = P.opts.alpha, lambda = P.opts.lambda, qFiltDecay = P.opts.q_filt_decay;
  int success = 1, rows = 1;
  double errOpt = 0.0;     // running maximum of the tracking error over all steps (err: over the steps before the first failure)
  const int lag = CTA ? (P.opts.early_exit ? 0 : AFFC_WORKERS - 1) : 0;
  if (CTA) { if (lane == 0) X->need = r_minDist; }

  // checkState + bookkeeping of step j (:730-752, :297-337), once its collision result is known.  pIk / pId0: the
  // failure the step found before looking at collisions; pEv / pEk: its tracking error if that grew, else -1.
  auto fold = [&](int j, const KCollResult& cmv, int pIk, int pId0, double pEv, int pEk) {
    int ikRes = pIk, id0 = pId0, id1 = -1;
    if (ikRes == 0 && cmv.minB1 >= 0 && cmv.minDist < 0.001) { ikRes = AFF_CODE_COLLISION; id0 = cmv.minB1; id1 = cmv.minB2; }
    if (ikRes != 0) {
      if (success) { r_first_code = ikRes; r_first_fail = j; }
      success = 0; r_code = ikRes; r_fail_step = j; r_id0 = id0; r_id1 = id1;
    }
    if (cmv.overflow) r_flags |= AFF_FLAG_LIST_OVERFLOW;
    collSum = collSum + cmv.cost;
    if (success && pEv > err) {
      err = pEv;
      if (err > AFFK_TRACK_MAX) {
        int etask = -1, edim = -1;
        for (int t = 0; t < nTasks; ++t) if (pEk >= c.tdesc[8 * t + TD_XOFF]) { etask = t; edim = pEk - c.tdesc[8 * t + TD_XOFF]; }
        success = 0; r_code = AFF_CODE_TRACKING; r_first_code = AFF_CODE_TRACKING;
        r_first_fail = j; r_fail_step = j; r_id0 = etask; r_id1 = edim;
      }
    }
    if (cmv.minDist < r_minDist) { r_minDist = cmv.minDist; r_b1 = cmv.minB1; r_b2 = cmv.minB2; }
  };

  // graph constraints due at time tAt (the start of a step): ConnectBodyConstraint, CollisionModelConstraint
  auto applyConstraints = [&](double tAt) {
  for (int k = 0; k < nGcon; ++k) {
    if ((gfired >> k) & 1u) continue;
    const double tc = ldg(&gcons[k].t);
    const int kind = ldg(&gcons[k].kind), slot = ldg(&gcons[k].slot);
    if (kind == AFF_CON_CONNECT) {
      if (tc - tAt < AFFK_TRAJ_EPS) {
        const int ps = ldg(&gcons[k].parent_slot);
        const int on = ldg(&P.obj_node[slot]);
        w_sync();
        if (lane < 12) {
          // HTr_invTransform: child relative to the new parent; or the identity the constraint
          // was given (setConnectTransform): the child then sits on the parent frame
          const double* Cc = k_slot(c, on); const double* Pp = k_slot(c, ps);
          double v;
          if (ldg(&gcons[k].on)) v = (lane == 3 || lane == 7 || lane == 11) ? 1.0 : 0.0;
          else if (lane < 3) {
            const double d0 = Cc[0] - Pp[0], d1 = Cc[1] - Pp[1], d2 = Cc[2] - Pp[2];
            v = fma(Pp[3 + 3 * lane + 2], d2, fma(Pp[3 + 3 * lane + 1], d1, Pp[3 + 3 * lane] * d0));
          } else {
            const int i = (lane - 3) / 3, j = (lane - 3) % 3;
            v = fma(Cc[3 + 3 * i + 2], Pp[3 + 3 * j + 2], fma(Cc[3 + 3 * i + 1], Pp[3 + 3 * j + 1], Cc[3 + 3 * i] * Pp[3 + 3 * j]));
          }
          sm[AFFK_O(P, objrel) + 12 * slot + lane] = v;
        }
        c.objHasRel |= 1u << slot;
        OBJ_SET(c, objParentSlot, slot, ps);
        OBJ_SET(c, objParentDynamic, slot, ps < P.nDyn);
        { const unsigned ch = c.chain[ps]; OBJ_SET(c, objChain, slot, ch); }
        const int pt = ldg(&gcons[k].parent_tree);
        { const int tr = (pt <= -2) ? OBJ_GET(c, objTree, -2 - pt) : pt; OBJ_SET(c, objTree, slot, tr); }
        w_sync();
        for (int i = lane; i < P.nDyn; i += 32) if (ldg(&P.dyn_under[i]) == slot + 1) c.chain[i] = ldg(&P.dyn_chain[i]) | OBJ_GET(c, objChain, slot);
        emask = ~(P.rbj_static_chain);
        emask &= ~(c.objChain0 | c.objChain1);
        gfired |= 1u << k;
      }
    } else {
      if (tc - tAt < 0.0) { const int md = ldg(&gcons[k].on) ? 1 : 2; OBJ_SET(c, objMode, slot, md); gfired |= 1u << k; }
    }
  }
  };
  if (CTA) { applyConstraints(0.0 + dt); w_post(&X->consSeq, 1); }

  K_MARK_INIT;
  for (int iter = 0; iter < nSteps; ++iter) {
    const int successPrev = success;
    const double tnext = tnow + dt;
    if (!CTA) {
      // ---- task switch / qFilt (:166-180)
      if (prevMask != activeMask) qFilt = 1.0;
      if (qFiltDecay <= 0.0) qFilt = 0.0;
      else qFilt = k_clip(qFilt - (1.0 / qFiltDecay) * dt, 0.0, 1.0);

      // ---- tc->step(dt): trajectories.  The state of an inactive task's trajectory is overwritten at
      // the start of the next step unless the task switches on in this one, so only then is it computed.
      const int ah = w_shfl(actHead, myTask < 0 ? 0 : myTask), ae = actEndOf(myTask < 0 ? 0 : myTask);
      // eight dimensions at a time: the solver scratch (via_stride doubles per lane) is the largest
      // user of the shared union region
      for (int batch = 0; batch < xdim; batch += AFFK_VIA_LANES) {
        if (lane >= batch && lane < batch + AFFK_VIA_LANES && lane < xdim) {
          const bool active = (activeMask >> myTask) & 1u;
          // does an activation point of my task fire in this step and switch it on?
          bool switchesOn = false;
          for (int k = ah; k < ae; ++k) { if (!(ldg(&acts[k].t) - tnext < AFFK_TRAJ_EPS)) break; switchesOn = ldg(&acts[k].on) != 0; }
          if (active || switchesOn) {
            double s3[3];
            if (!active) { s3[0] = xcur[lane]; s3[1] = 0.0; s3[2] = 0.0; }
            else { s3[0] = st[3 * lane]; s3[1] = st[3 * lane + 1]; s3[2] = st[3 * lane + 2]; }
            k_via_step(s3, vias + viaHead, viaEnd - viaHead, tnow, dt, sm + AFFK_O(P, scratch) + P.via_stride * (lane - batch), P.via_unk);
            st[3 * lane] = s3[0]; st[3 * lane + 1] = s3[1]; st[3 * lane + 2] = s3[2];
            xdes[lane] = s3[0];
          }
        }
        w_sync();
      }
    }
    K_MARK(0);
    tnow = tnext;
    // graph constraints (block-per-candidate: applied at the end of the step before, so that the null-space
    // warp can start on the state they leave)
    if (!CTA) applyConstraints(tnow);
    double alphaEff, phaseScale;
    const double* xdesNow = xdes;
    if (!CTA) {
      // activation points
      int myActive = (activeMask >> lane) & 1u;
      if (lane < nTasks) {
        const int actEnd = actEndOf(lane);
        while (actHead < actEnd && ldg(&acts[actHead].t) - tnow < AFFK_TRAJ_EPS) { myActive = ldg(&acts[actHead].on); ++actHead; }
      }
      if (lane < xdim) { const int ve = viaEnd; while (viaHead < ve && ldg(&vias[viaHead].t) - tnow < AFFK_TRAJ_EPS) ++viaHead; }
      prevMask = activeMask;
      activeMask = w_ballot(lane < nTasks && myActive);
      {
        const double rem = endAbs - tnow;
        endTime = rem > 0.0 ? rem : 0.0;
      }
      // blending
      double bl = 1.0;
      if (lane < nTasks) {
        if (actHead < actEndOf(lane)) { const double h = ldg(&acts[actHead].horizon); if (h > 0.0) bl = k_clip((ldg(&acts[actHead].t) - tnow) / h, 0.0, 1.0); }
        if (actHead > actBeginOf(lane)) {
          const double lastSwitchT = ldg(&acts[actHead - 1].t), lastSwitchH = ldg(&acts[actHead - 1].horizon);
          if (lastSwitchH > 0.0) { const double up = k_clip((tnow - lastSwitchT) / lastSwitchH, 0.0, 1.0); if (up < bl) bl = up; }
        }
      }
      double blending = w_ballot(bl < 1.0) ? w_min(bl) : 1.0;
      if (activeMask == 0u) blending = 0.0;
      const double phase = (motionDuration > 0.0) ? 1.0 - (endTime / motionDuration) : 0.0;
      phaseScale = aff_sin(AFF_PI * phase);
      alphaEff = blending * alpha;
      w_sync();

    } else {
      // the record of this step from the trajectory warp
      w_wait_ge(&X->trajSeq, iter + 1);
      const KCtaStep* rec = &X->ring[iter & (AFFC_TRING - 1)];
      activeMask = rec->activeMask; alphaEff = rec->alphaEff; phaseScale = rec->phaseScale; qFilt = rec->qFilt;
      xdesNow = rec->xdes;
    }
    K_MARK(1);
    // ================= computeIK =================
    // row offsets of the active tasks
    int m = 0, myRow = 0;
    for (int t = 0; t < nTasks; ++t) {
      if (t == lane) myRow = m;
      if ((activeMask >> t) & 1u) m += c.tdesc[8 * t + TD_DIM];
    }
    if (lane < nTasks && ((activeMask >> lane) & 1u)) k_task_dx(c.tdesc + 8 * lane, tasks + lane, geo + AFFK_GEO * lane, xcur, xdesNow, dxv + myRow);
    // joint frame of this lane's IK joint
    KJointFrame jf;
    jf.valid = (lane < nJ); jf.type = jtype;
    {
      const double* A = k_slot(c, jnode);
      const int ar = (jtype < AFF_JNT_ROT_X) ? jtype : jtype - AFF_JNT_ROT_X;
      for (int k = 0; k < 3; ++k) { jf.ax[k] = A[3 + 3 * ar + k]; jf.org[k] = A[k]; }
    }
    // J columns; nz = rows of this column that are non-zero
    unsigned nz = 0u;
    {
      int row = 0;
      for (int t = 0; t < nTasks; ++t) {
        if (!((activeMask >> t) & 1u)) continue;
        if (lane < nJ) nz |= k_task_J(c, jf, lane, c.tdesc + 8 * t, tasks + t, geo + AFFK_GEO * t, J + row * JS, JS) << row;
        row += c.tdesc[8 * t + TD_DIM];
      }
    }
    K_MARK(2);
    // dH: joint limit gradient + elbow / wrist terms (:589-625); block-per-candidate: from the null-space warp
    double dHv;
    if (!CTA) dHv = k_null_space_dH(c, jf, q, alphaEff, phaseScale);
    else { w_wait_ge(&X->nsSeq, iter + 1); dHv = X->dH[lane]; }
    // joint masks (:636-675); row masks of J for the sparse products below
    {
      const unsigned aMask = w_ballot(nz != 0u);
      jmaskBits |= aMask;
      if (lane < nJ && !(((aMask | emask) >> lane) & 1u)) dHv = dHv * 0.0;
      for (int r = 0; r < m; ++r) { const unsigned rm = w_ballot((nz >> r) & 1u); if (lane == r) rmask[r] = rm; }
    }
    K_MARK(3);
    // ---- solveRightInverse
    const double vv = (lane < nJ) ? jw * dHv : 0.0;
    if (lane < nJ) vbuf[lane] = vv;          // v = W^-1 dH
    w_sync();
    if (lane < m) {
      double acc = dxv[lane];
      unsigned mk = rmask[lane];
      while (mk) { const int j = w_ffs(mk) - 1; mk &= mk - 1u; acc = fma(J[lane * JS + j], vbuf[j], acc); }
      rhs[lane] = acc;
    }
    {
      const int nEnt = m * (m + 1) / 2;
      for (int ebase = 0; ebase < nEnt; ebase += 32) {
        const int e = ebase + lane;
        if (e < nEnt) {
          // e -> (i, k), k <= i
          int i = (int)((sqrtf(8.0f * (float)e + 1.0f) - 1.0f) * 0.5f);
          while ((i + 1) * (i + 2) / 2 <= e) ++i;
          while (i * (i + 1) / 2 > e) --i;
          const int k = e - i * (i + 1) / 2;
          double acc = 0.0;
          unsigned mk = rmask[i] & rmask[k];
          while (mk) { const int j = w_ffs(mk) - 1; mk &= mk - 1u; acc = fma(J[i * JS + j] * wbuf[j], J[k * JS + j], acc); }
          if (i == k) acc = acc + lambda;
          L[i * LS + k] = acc;
        }
      }
    }
    w_sync();
    K_MARK(4);
    // Cholesky, one lane per row, reciprocal pivots
    int singular = 0;
    double myInv = 1.0;     // lane i keeps 1 / L[i][i]
    for (int j = 0; j < m; ++j) {
      double t = 0.0;
      if (lane >= j && lane < m) {
        t = L[lane * LS + j];
#pragma unroll 4
        for (int k = 0; k < j; ++k) t = fma(-L[lane * LS + k], L[j * LS + k], t);
      }
      const double sdiag = w_shfl(t, j);
      if (!(sdiag > 0.0)) { singular = 1; break; }
      const double ljj = sqrt(sdiag);
      const double inv = 1.0 / ljj;
      if (lane == j) { L[j * LS + j] = ljj; myInv = inv; }
      else if (lane > j && lane < m) L[lane * LS + j] = t * inv;
      w_sync();
    }
    K_MARK(5);
    int ikRes = 0, id0 = -1;      // failure of this step, collisions aside
    // block-per-candidate: the worker of the previous step has taken that step's frames and object state
    if (CTA) { if (iter > 0) w_wait_ge(&X->copied[(iter - 1) % AFFC_WORKERS], iter); }
    if (singular) {
      ikRes = AFF_CODE_SINGULAR;
    } else {
      // forward substitution L y = rhs (lane i owns y[i]); column oriented
      double yi = (lane < m) ? rhs[lane] : 0.0;
      for (int i = 0; i < m; ++i) {
        if (lane == i) yi = yi * myInv;
        const double yfin = w_shfl(yi, i);
        if (lane > i && lane < m) yi = fma(-L[lane * LS + i], yfin, yi);
      }
      // backward substitution L^T z = y
      for (int i = m - 1; i >= 0; --i) {
        if (lane == i) yi = yi * myInv;
        const double zfin = w_shfl(yi, i);
        if (lane < i) yi = fma(-L[i * LS + lane], zfin, yi);
      }
      if (lane < m) rhs[lane] = yi;
      w_sync();
      double dqv = 0.0;
      if (lane < nJ) {
        double acc = 0.0;
        unsigned mk = nz;
        while (mk) { const int i = w_ffs(mk) - 1; mk &= mk - 1u; acc = fma(J[i * JS + lane], rhs[i], acc); }
        dqv = fma(jw, acc, -vv);
      }
      // speed and acceleration limits (:696-723)
      {
        double sc = 1.0;
        if (lane < nJ && jspeed < AFFK_NO_LIMIT) { const double a = fabs(dqv), mx = jspeed * dt; if (a > mx) sc = mx / a; }
        if (w_ballot(sc < 1.0)) { const double scale = w_min(sc); dqv = dqv * (0.99999 * scale); }
      }
      const double dynLim = -0.99 * qFilt + 1.0;
      const double invdt = 1.0 / dt;
      double qd = 0.0;
      if (lane < nJ) {
        qd = dqv * invdt;
        const double qdPrev = qdot[lane];
        if (jacc < AFFK_NO_LIMIT) {
          const double lim = (dynLim * jacc) * dt;
          const double dv = qd - qdPrev;
          if (dv > lim) qd = qdPrev + lim;
          else if (dv < -lim) qd = qdPrev - lim;
        }
        qdot[lane] = qd; q[lane] = q[lane] + dqv;     // own elements only: no hazard
      }
      w_sync();
      K_MARK(6);
      // integration, FK, collision model, checkState (:730-752); block-per-candidate: the rounds with the frames the
      // step loop reads, the FK warp does the rest
      k_fk_t<CTA>(c, false, 0, CTA ? P.nRoundsCrit : P.nRounds);
      K_MARK(7);
      if (!CTA) cm = k_collision(c, r_minDist > 0.001 ? r_minDist : 0.001);
      if (!CTA) {     // (block-per-candidate: the null-space warp checks the new state and adds up its joint-limit cost)
        const unsigned fast = w_ballot(lane < nJ && jspeed < AFFK_NO_LIMIT && fabs(qd) > jspeed * 1.0);
        const unsigned viol = w_ballot(lane < nJ && (q[lane] < jqmin || q[lane] > jqmax));
        if (fast) ikRes = AFF_CODE_SPEED;
        else if (viol) { ikRes = AFF_CODE_JOINT_LIMIT; id0 = w_popc(viol); }
      }
    }
    if (CTA) {
      // hand the step to its worker (a singular step leaves the frames as they were: the worker finds what
      // the single-warp kernel keeps from the step before)
      if (lane == 0) { X->objState[0] = c.objMode0; X->objState[1] = c.objMode1; X->objState[2] = c.objTree0; X->objState[3] = c.objTree1; }
      w_post(&X->fkSeq, iter + 1);
      if (iter + 1 < nSteps) applyConstraints(tnow + dt);
      w_post(&X->consSeq, iter + 2);
    }
    // ---- costs (:226-227)
    if (!CTA) {
      double term = 0.0;
      if (lane < nJ) {
        const double dqc = q[lane] - jq0;
        const double r = (dqc > 0.0) ? jqmax - jq0 : jq0 - jqmin;
        const double u = dqc / r;
        term = (0.5 * jwjl) * (u * u);
      }
      jlSum = jlSum + w_tree_sum(term);
    }
    K_MARK(8);
    // ---- geometry of the new state; tracking error (:297-337)
    const long long tgeo_ = K_NOW();
    if (CTA) { if (iter > 0) w_wait_ge(&X->trkSeq, iter); }      // the tracking error of the step before has been taken from geo / xcur
    // block-per-candidate: the orientation tasks (acos / atan2 / normalisations: the long branches) are the geometry warp's
    {
      const bool orient = lane < nTasks && k_geo_helper_kind(c.tdesc[8 * (lane < nTasks ? lane : 0) + TD_KIND]);
      if (lane < nTasks && !orient) k_task_geo(c, c.tdesc + 8 * lane, tasks + lane, geo + AFFK_GEO * lane, xcur);
      // one warp: the orientation tasks share one normalisation and one acos instead of a branch each
      if (!CTA) k_task_geo_orient(c, c.tdesc + 8 * (orient ? lane : 0), geo + AFFK_GEO * (orient ? lane : 0), xcur, orient);
    }
    if (CTA) w_wait_ge(&X->geoHelpSeq, iter + 1);
    w_sync();
    K_SUB(12, tgeo_);
    if (CTA) w_post(&X->geoSeq, iter + 2);      // task values of this state are final
    double pEv = -1.0; int pEk = 0;
    if (!CTA && success && ikRes == 0) {       // (a step that has failed already does not look at the tracking error)
      double* dxe = dxv;   // stacked over all dims here
      if (lane < nTasks && ((activeMask >> lane) & 1u)) k_task_dx(c.tdesc + 8 * lane, tasks + lane, geo + AFFK_GEO * lane, xcur, xdesNow, dxe + c.tdesc[8 * lane + TD_XOFF]);
      w_sync();
      double ev = -1.0; int ek = 0x7fffffff;
      if (lane < xdim && ((activeMask >> myTask) & 1u)) { ev = fabs(dxe[lane]); ek = lane; }
      if (w_ballot(ev > errOpt)) {        // the running maximum grows
        w_max_keyed(ev, ek);
        errOpt = ev; pEv = ev; pEk = ek;
      }
      w_sync();
    }
    K_MARK(9);
    if (!CTA) fold(iter, cm, ikRes, id0, pEv, pEk);
    else {
      const int slot = iter & (AFFC_RING - 1);
      if (lane == 0) { X->pendIk[slot] = ikRes; X->pendId0[slot] = id0; }     // (pendEv / pendEk: from the null-space warp)
      w_sync();
      const int j = iter - lag;
      if (j >= 0) {
        w_wait_ge(&X->done[j % AFFC_WORKERS], j + 1);
        w_wait_ge(&X->trkSeq, j + 1);
        const int sj = j & (AFFC_RING - 1);
        const KCollResult cj = X->res[sj];
        const int sing = X->pendIk[sj];
        fold(j, cj, sing ? sing : X->pendChk[sj], sing ? -1 : X->pendChkId0[sj], X->pendEv[sj], X->pendEk[sj]);
        jlSum = X->pendJl[sj];
        if (lane == 0) X->need = r_minDist;
      }
    }
    K_MARK(10);
    if (qlog && lane < nJ && rows < qrows) qlog[(size_t)rows * nJ + lane] = q[lane];
    rows++;
    if (P.opts.early_exit && !success && successPrev) break;     // opt-in: the reference's commented-out break (:368-373)
  }
  if (CTA) {
    // the steps still in flight
    const int stepsDone = rows - 1;
    for (int j = (stepsDone > lag ? stepsDone - lag : 0); j < stepsDone; ++j) {
      w_wait_ge(&X->done[j % AFFC_WORKERS], j + 1);
      w_wait_ge(&X->trkSeq, j + 1);
      const int sj = j & (AFFC_RING - 1);
      const KCollResult cj = X->res[sj];
      const int sing = X->pendIk[sj];
      fold(j, cj, sing ? sing : X->pendChk[sj], sing ? -1 : X->pendChkId0[sj], X->pendEv[sj], X->pendEk[sj]);
      jlSum = X->pendJl[sj];
    }
    w_post(&X->stop, 1);
  }
  if (lane == 0) {
    aff_result res;
    res.idx = cand; res.success = success; res.code = r_code; res.first_code = r_first_code;
    res.first_fail_step = r_first_fail; res.fail_step = r_fail_step; res.fail_id0 = r_id0; res.fail_id1 = r_id1;
    res.min_dist_b1 = r_b1; res.min_dist_b2 = r_b2; res.n_steps = rows; res.reserved = r_flags;
    res.jmask_bits = (uint64_t)jmaskBits; res.min_dist = r_minDist;
    res.jl_cost = jlSum / (double)rows; res.coll_cost = collSum / (double)rows; res.track_err = err;
    P.results[cand] = res;
  }
}

AFF_DEV void k_rollout(const KPlan& P, int cand, double* sm) { k_rollout_t<0>(P, cand, sm, nullptr); }
AFF_DEV int k_xcur_offset(const KPlan& P) { return AFFK_O(P, xcur); }

// Worker warp w (1 .. AFFC_WORKERS) of a block-per-candidate launch: the collision model of the steps
// k = w - 1, w - 1 + AFFC_WORKERS, ... on a private copy of the frames (its own slab: same layout as warp 0's).
AFF_DEV void k_cta_worker(const KPlan& P, int cand, double* sm, const double* nodes0, KCtaShared* X, int w)
{
  KCtx c;
  c.P = &P; c.C = P.cands + cand; c.sm = sm; c.lane = w_lane();
  c.nodes = sm + P.o_node;
  c.chain = (unsigned*)(sm + AFFK_O(P, chain));
  c.tdesc = (int*)(sm + AFFK_O(P, tdesc));
  c.objHasRel = 0u;
  c.objParentSlot0 = c.objParentSlot1 = 0; c.objParentDynamic0 = c.objParentDynamic1 = 0; c.objChain0 = c.objChain1 = 0u;
  c.objTree0 = c.objTree1 = -1; c.objMode0 = c.objMode1 = 0;
  const int lane = c.lane;
  const int nEl = 12 * (P.nDyn + P.nStatPar);
  const int nSteps = ldg(&c.C->n_steps);
  for (int k = w - 1; k < nSteps; k += AFFC_WORKERS) {
    if (w_wait_ge_or(&X->fk2Seq, k + 1, &X->stop)) return;
    for (int i = lane; i < nEl; i += 32) c.nodes[i] = nodes0[i];
    c.objMode0 = X->objState[0]; c.objMode1 = X->objState[1]; c.objTree0 = X->objState[2]; c.objTree1 = X->objState[3];
    w_post(&X->copied[w - 1], k + 1);
    const double run = c_ld(&X->need);
    const KCollResult cm = k_collision(c, run > 0.001 ? run : 0.001);
    if (lane == 0) X->res[k & (AFFC_RING - 1)] = cm;
    w_post(&X->done[w - 1], k + 1);
  }
}
// Trajectory warp of a block-per-candidate launch: everything of a step that is driven by time alone
// (:166-180 qFilt, tc->step, activation points, blending), up to AFFC_TRING steps ahead of the step loop.
// It waits for the rollout only where a task switches on: that trajectory starts from the task value of
// the state before the step (xcur0: warp 0's slab).
AFF_DEV void k_cta_traj(const KPlan& P, int cand, double* sm, const double* xcur0, KCtaShared* X)
{
  const KCand& C = P.cands[cand];
  const int lane = w_lane();
  const double dt = P.opts.dt;
  const int nTasks = ldg(&C.n_tasks), xdim = ldg(&C.xdim);
  const int taskBase = ldg(&C.first_task);
  double* st = sm + AFFK_O(P, st); double* xdes = sm + AFFK_O(P, xdes);
  int myTask = -1;
  if (lane < xdim) for (int t = 0; t < nTasks; ++t) if (lane >= ldg(&tasks[t].xoff)) myTask = t;
  int viaHead = 0;
  if (lane < xdim) viaHead = ldg(&C.first_via) + ldg(&C.via_off[lane]);
  int actHead = 0;
  if (lane < nTasks) actHead = ldg(&C.first_act) + ldg(&C.act_off[lane]);
  if (lane < xdim) { st[3 * lane] = 0.0; st[3 * lane + 1] = 0.0; st[3 * lane + 2] = 0.0; xdes[lane] = 0.0; }
  w_sync();
  unsigned activeMask = 0u, prevMask = 0u;
  double tnow = 0.0, qFilt = 0.0, endTime = ldg(&C.end_abs);
  const double endAbs = ldg(&C.end_abs);
  const double motionDuration = endAbs - ldg(&C.start_abs);
  const double alpha = P.opts.alpha, qFiltDecay = P.opts.q_filt_decay;
  const int nSteps = ldg(&C.n_steps);
  for (int iter = 0; iter < nSteps; ++iter) {
    // the ring slot of this step is free once the last reader (tracking error) of step iter - AFFC_TRING is through
    if (iter >= AFFC_TRING) { if (w_wait_ge_or(&X->trkSeq, iter - AFFC_TRING + 1, &X->stop)) return; }
    if (prevMask != activeMask) qFilt = 1.0;
    if (qFiltDecay <= 0.0) qFilt = 0.0;
    else qFilt = k_clip(qFilt - (1.0 / qFiltDecay) * dt, 0.0, 1.0);
    const double tnext = tnow + dt;
    const int ah = w_shfl(actHead, myTask < 0 ? 0 : myTask), ae = actEndOf(myTask < 0 ? 0 : myTask);
    bool active = false, switchesOn = false;
    if (lane < xdim) {
      active = (activeMask >> myTask) & 1u;
      for (int k = ah; k < ae; ++k) { if (!(ldg(&acts[k].t) - tnext < AFFK_TRAJ_EPS)) break; switchesOn = ldg(&acts[k].on) != 0; }
    }
    if (w_ballot(!active && switchesOn)) { if (w_wait_ge_or(&X->geoSeq, iter + 1, &X->stop)) return; }
    for (int batch = 0; batch < xdim; batch += AFFK_VIA_LANES) {
      if (lane >= batch && lane < batch + AFFK_VIA_LANES && lane < xdim) {
        if (active || switchesOn) {
          double s3[3];
          if (!active) { s3[0] = xcur0[lane]; s3[1] = 0.0; s3[2] = 0.0; }
          else { s3[0] = st[3 * lane]; s3[1] = st[3 * lane + 1]; s3[2] = st[3 * lane + 2]; }
          k_via_step(s3, vias + viaHead, viaEnd - viaHead, tnow, dt, sm + AFFK_O(P, scratch) + P.via_stride * (lane - batch), P.via_unk);
          st[3 * lane] = s3[0]; st[3 * lane + 1] = s3[1]; st[3 * lane + 2] = s3[2];
          xdes[lane] = s3[0];
        }
      }
      w_sync();
    }
    tnow = tnext;
    int myActive = (activeMask >> lane) & 1u;
    if (lane < nTasks) {
      const int actEnd = actEndOf(lane);
      while (actHead < actEnd && ldg(&acts[actHead].t) - tnow < AFFK_TRAJ_EPS) { myActive = ldg(&acts[actHead].on); ++actHead; }
    }
    if (lane < xdim) { const int ve = viaEnd; while (viaHead < ve && ldg(&vias[viaHead].t) - tnow < AFFK_TRAJ_EPS) ++viaHead; }
    prevMask = activeMask;
    activeMask = w_ballot(lane < nTasks && myActive);
    {
      const double rem = endAbs - tnow;
      endTime = rem > 0.0 ? rem : 0.0;
    }
    double bl = 1.0;
    if (lane < nTasks) {
      if (actHead < actEndOf(lane)) { const double h = ldg(&acts[actHead].horizon); if (h > 0.0) bl = k_clip((ldg(&acts[actHead].t) - tnow) / h, 0.0, 1.0); }
      if (actHead > actBeginOf(lane)) {
        const double lastSwitchT = ldg(&acts[actHead - 1].t), lastSwitchH = ldg(&acts[actHead - 1].horizon);
        if (lastSwitchH > 0.0) { const double up = k_clip((tnow - lastSwitchT) / lastSwitchH, 0.0, 1.0); if (up < bl) bl = up; }
      }
    }
    double blending = w_ballot(bl < 1.0) ? w_min(bl) : 1.0;
    if (activeMask == 0u) blending = 0.0;
    const double phase = (motionDuration > 0.0) ? 1.0 - (endTime / motionDuration) : 0.0;
    KCtaStep* rec = &X->ring[iter & (AFFC_TRING - 1)];
    if (lane < xdim) rec->xdes[lane] = xdes[lane];
    if (lane == 0) { rec->activeMask = activeMask; rec->alphaEff = blending * alpha; rec->phaseScale = aff_sin(AFF_PI * phase); rec->qFilt = qFilt; }
    w_post(&X->trajSeq, iter + 1);
  }
}

// FK warp of a block-per-candidate launch: the rounds [nRoundsCrit, nRounds) of the FK schedule (frames only the
// collision model reads: fingers, parts of the robot no task refers to), in warp 0's node array, once warp 0
// has the early rounds of the step and the collision worker of the step before has taken its copy.
AFF_DEV void k_cta_fk(const KPlan& P, int cand, double* sm0, KCtaShared* X)
{
  KCtx c;
  c.P = &P; c.C = P.cands + cand; c.sm = sm0; c.lane = w_lane();
  c.nodes = sm0 + P.o_node;
  c.chain = (unsigned*)(sm0 + AFFK_O(P, chain));
  c.tdesc = (int*)(sm0 + AFFK_O(P, tdesc));
  c.objHasRel = 0u;      // (objects and what hangs below them are in the early rounds)
  c.objParentSlot0 = c.objParentSlot1 = 0; c.objParentDynamic0 = c.objParentDynamic1 = 0; c.objChain0 = c.objChain1 = 0u;
  c.objTree0 = c.objTree1 = -1; c.objMode0 = c.objMode1 = 0;
  const int nSteps = ldg(&c.C->n_steps);
  for (int iter = 0; iter < nSteps; ++iter) {
    if (w_wait_ge_or(&X->fkSeq, iter + 1, &X->stop)) return;
    if (iter > 0) { if (w_wait_ge_or(&X->copied[(iter - 1) % AFFC_WORKERS], iter, &X->stop)) return; }
    k_fk_t<1>(c, false, P.nRoundsCrit, P.nRounds);
    w_post(&X->fk2Seq, iter + 1);
  }
}

// Geometry warp of a block-per-candidate launch: the task values of the orientation tasks of the state after
// step k (frames final once fkSeq > k; geo / xcur rows of the step before free once trkSeq >= k), written into
// warp 0's slab next to the rows warp 0 computes meanwhile.
AFF_DEV void k_cta_geometry(const KPlan& P, int cand, double* sm0, KCtaShared* X)
{
  KCtx c;
  c.P = &P; c.C = P.cands + cand; c.sm = sm0; c.lane = w_lane();
  c.nodes = sm0 + P.o_node;
  c.chain = (unsigned*)(sm0 + AFFK_O(P, chain));
  c.tdesc = (int*)(sm0 + AFFK_O(P, tdesc));
  c.objHasRel = 0u;
  c.objParentSlot0 = c.objParentSlot1 = 0; c.objParentDynamic0 = c.objParentDynamic1 = 0; c.objChain0 = c.objChain1 = 0u;
  c.objTree0 = c.objTree1 = -1; c.objMode0 = c.objMode1 = 0;
  const KCand& C = *c.C;
  const int lane = c.lane;
  const int nTasks = ldg(&C.n_tasks);
  const int taskBase = ldg(&C.first_task);
  double* geo = sm0 + AFFK_O(P, geo); double* xcur = sm0 + AFFK_O(P, xcur);
  const int nSteps = ldg(&C.n_steps);
  for (int iter = 0; iter < nSteps; ++iter) {
    if (w_wait_ge_or(&X->fkSeq, iter + 1, &X->stop)) return;
    if (iter > 0) { if (w_wait_ge_or(&X->trkSeq, iter, &X->stop)) return; }
    {
      const bool on = lane < nTasks && k_geo_helper_kind(c.tdesc[8 * (lane < nTasks ? lane : 0) + TD_KIND]);
      k_task_geo_orient(c, c.tdesc + 8 * (on ? lane : 0), geo + AFFK_GEO * (on ? lane : 0), xcur, on);
    }
    w_post(&X->geoHelpSeq, iter + 1);
  }
}

// Null-space warp of a block-per-candidate launch: the two per-step computations that sit beside the
// critical chain FK -> Jacobian -> solve -> FK.  (1) dH of step k from the state before it (frames, chains,
// q in warp 0's slab: final once consSeq > k) and the step's blending / phase factors; (2) the tracking
// error of the state after step k (:297-337; geo / xcur final once geoSeq > k + 1), unconditionally: the
// step loop decides at fold time whether it still counts.
AFF_DEV void k_cta_nullspace(const KPlan& P, int cand, double* sm0, KCtaShared* X)
{
  KCtx c;
  c.P = &P; c.C = P.cands + cand; c.sm = sm0; c.lane = w_lane();
  c.nodes = sm0 + P.o_node;
  c.chain = (unsigned*)(sm0 + AFFK_O(P, chain));
  c.tdesc = (int*)(sm0 + AFFK_O(P, tdesc));
  c.objHasRel = 0u;
  c.objParentSlot0 = c.objParentSlot1 = 0; c.objParentDynamic0 = c.objParentDynamic1 = 0; c.objChain0 = c.objChain1 = 0u;
  c.objTree0 = c.objTree1 = -1; c.objMode0 = c.objMode1 = 0;
  const KCand& C = *c.C;
  const int lane = c.lane, nJ = P.nJ;
  const int nTasks = ldg(&C.n_tasks), xdim = ldg(&C.xdim);
  const int taskBase = ldg(&C.first_task);
  const double* q = sm0 + AFFK_O(P, q);
  const double* geo = sm0 + AFFK_O(P, geo); const double* xcur = sm0 + AFFK_O(P, xcur);
  double* dxe = X->trkDx;
  int jnode = 0, jtype = 0;
  if (lane < nJ) { jnode = ldg(&P.j_node[lane]); jtype = ldg(&P.j_type[lane]); }
  int myTask = -1;
  if (lane < xdim) for (int t = 0; t < nTasks; ++t) if (lane >= ldg(&tasks[t].xoff)) myTask = t;
  const int nSteps = ldg(&C.n_steps);
  // tracking error of the state after step j
  auto tracking = [&](int j) -> bool {
    if (w_wait_ge_or(&X->geoSeq, j + 2, &X->stop)) return true;
    const KCtaStep* rec = &X->ring[j & (AFFC_TRING - 1)];
    const unsigned activeMask = rec->activeMask;
    if (lane < nTasks && ((activeMask >> lane) & 1u)) k_task_dx(c.tdesc + 8 * lane, tasks + lane, geo + AFFK_GEO * lane, xcur, rec->xdes, dxe + c.tdesc[8 * lane + TD_XOFF]);
    w_sync();
    double ev = -1.0; int ek = 0x7fffffff;
    if (lane < xdim && ((activeMask >> myTask) & 1u)) { ev = fabs(dxe[lane]); ek = lane; }
    w_max_keyed(ev, ek);
    const int slot = j & (AFFC_RING - 1);
    if (lane == 0) { X->pendEv[slot] = ev; X->pendEk[slot] = ek; }
    w_post(&X->trkSeq, j + 1);
    return false;
  };
  // checkState of the state after step j, collisions aside (:756-886 speed, joint limits), and its joint-limit cost (:226-227)
  const double* qdot = sm0 + AFFK_O(P, qdot);
  const int jl = (lane < nJ) ? lane : 0;
  double jlSum = 0.0;
  auto checks = [&](int j, double qv, double qdv) {
    // (jq0, jqmin, ... : the per-joint constants of this lane, re-read through the macros of k_rollout_t)
    const unsigned fast = w_ballot(lane < nJ && jspeed < AFFK_NO_LIMIT && fabs(qdv) > jspeed * 1.0);
    const unsigned viol = w_ballot(lane < nJ && (qv < jqmin || qv > jqmax));
    double term = 0.0;
    if (lane < nJ) {
      const double dqc = qv - jq0;
      const double r = (dqc > 0.0) ? jqmax - jq0 : jq0 - jqmin;
      const double u = dqc / r;
      term = (0.5 * jwjl) * (u * u);
    }
    jlSum = jlSum + w_tree_sum(term);
    const int slot = j & (AFFC_RING - 1);
    if (lane == 0) {
      X->pendChk[slot] = fast ? AFF_CODE_SPEED : (viol ? AFF_CODE_JOINT_LIMIT : 0);
      X->pendChkId0[slot] = (!fast && viol) ? w_popc(viol) : -1;
      X->pendJl[slot] = jlSum;
    }
  };
  for (int iter = 0; iter < nSteps; ++iter) {
    // dH of step iter first (the step loop needs it midway through the step), then the checks and the tracking error
    // of the step before.  q / qdot of the state before step iter are taken before dH is posted: the step loop
    // overwrites them once it has dH.
    if (w_wait_ge_or(&X->consSeq, iter + 1, &X->stop)) return;
    if (w_wait_ge_or(&X->trajSeq, iter + 1, &X->stop)) return;
    double qv = 0.0, qdv = 0.0;
    if (lane < nJ) { qv = q[lane]; qdv = qdot[lane]; }
    {
      const KCtaStep* rec = &X->ring[iter & (AFFC_TRING - 1)];
      KJointFrame jf;
      jf.valid = (lane < nJ); jf.type = jtype;
      const double* A = k_slot(c, jnode);
      const int ar = (jtype < AFF_JNT_ROT_X) ? jtype : jtype - AFF_JNT_ROT_X;
      for (int k = 0; k < 3; ++k) { jf.ax[k] = A[3 + 3 * ar + k]; jf.org[k] = A[k]; }
      const double dHv = k_null_space_dH(c, jf, q, rec->alphaEff, rec->phaseScale);
      X->dH[lane] = dHv;
      w_post(&X->nsSeq, iter + 1);
    }
    if (iter > 0) { checks(iter - 1, qv, qdv); if (tracking(iter - 1)) return; }
  }
  if (nSteps > 0) {
    if (w_wait_ge_or(&X->fkSeq, nSteps, &X->stop)) return;
    double qv = 0.0, qdv = 0.0;
    if (lane < nJ) { qv = q[lane]; qdv = qdot[lane]; }
    checks(nSteps - 1, qv, qdv);
    tracking(nSteps - 1);
  }
}

#undef viaEnd
#undef actBeginOf
#undef actEndOf
#undef vias
#undef acts
#undef gcons
#undef tasks
#undef jq0
#undef jqmin
#undef jqmax
#undef jspeed
#undef jacc
#undef jwjl
#undef jw

// ------------------------------------------------- prologue (once per batch)
// World transforms of the static nodes, world frames / AABBs of the static
// shapes, and sin/cos of the constant joints.  One warp.
AFF_DEV void k_prologue(const KPlan& P)
{
  const int lane = w_lane();
  // sin / cos of constant joints
  for (int n = lane; n < P.nStatic; n += 32) {
    double s = 0.0, co = 1.0;
    if (ldg(&P.stat[n].kind) == AFFK_NODE_JOINT_CONST && ldg(&P.stat[n].jtype) >= AFF_JNT_ROT_X) aff_sincos(ldg(&P.stat[n].qconst), &s, &co);
    P.statSC[2 * n] = s; P.statSC[2 * n + 1] = co;
  }
  for (int n = lane; n < P.nDyn; n += 32) {
    double s = 0.0, co = 1.0;
    if (ldg(&P.dyn[n].kind) == AFFK_NODE_JOINT_CONST && ldg(&P.dyn[n].jtype) >= AFF_JNT_ROT_X) aff_sincos(ldg(&P.dyn[n].qconst), &s, &co);
    P.dynSC[2 * n] = s; P.dynSC[2 * n + 1] = co;
  }
  if (lane < 12) P.statA[12 * P.nStatic + lane] = (lane == 3 || lane == 7 || lane == 11) ? 1.0 : 0.0;
  w_sync();
  // static nodes, parents first; element-parallel like k_fk
  const int e = lane % 12;
  for (int n = 0; n < P.nStatic; ++n) {
    const KNode* nd = P.stat + n;
    const int par = nd->parent;
    const double* Pp = (par == AFFK_WORLD) ? P.statA + 12 * P.nStatic : P.statA + 12 * AFFK_STATIC_IDX(par);
    double val = nd->identityT ? Pp[e] : k_compose_elem(nd->T, Pp, e);
    int partner = lane, isA = 0, isB = 0;
    const int kind = nd->kind, jt = nd->jtype;
    if (kind == AFFK_NODE_JOINT_CONST) {
      if (jt >= AFF_JNT_ROT_X && e >= 3) {
        const int d = jt - AFF_JNT_ROT_X, a = (d + 1) % 3, b = (d + 2) % 3, row = (e - 3) / 3, col = (e - 3) % 3;
        if (row == a) { isA = 1; partner = 3 + 3 * b + col; } else if (row == b) { isB = 1; partner = 3 + 3 * a + col; }
      } else if (jt < AFF_JNT_ROT_X && e < 3) partner = 3 + 3 * jt + e;
    }
    const double other = w_shfl(val, partner);
    if (kind == AFFK_NODE_JOINT_CONST) {
      const double s = P.statSC[2 * n], co = P.statSC[2 * n + 1];
      if (jt >= AFF_JNT_ROT_X) { if (isA) val = fma(s, other, co * val); else if (isB) val = fma(co, val, -(s * other)); }
      else if (e < 3) val = fma(nd->qconst, other, val);
    }
    if (lane < 12) P.statA[12 * n + e] = val;
    w_sync();
  }
  // static shapes
  for (int s = lane; s < P.nStatShapes; s += 32) {
    const KShape* sh = P.sshape + s;
    const int ref = sh->node;
    const double* An = (ref == AFFK_WORLD) ? P.statA + 12 * P.nStatic : P.statA + 12 * AFFK_STATIC_IDX(ref);
    double W[12];
    k_compose(W, sh->A_CB, An);
    for (int k = 0; k < 12; ++k) P.sshapeA[12 * s + k] = W[k];
    // segment record of capsules / spheres: p, d = length * z-axis, r
    const bool ssl = sh->type == AFF_SHAPE_SSL;
    for (int k = 0; k < 3; ++k) { P.sseg[8 * s + k] = W[k]; P.sseg[8 * s + 3 + k] = ssl ? sh->ext[2] * W[9 + k] : 0.0; }
    P.sseg[8 * s + 6] = sh->ext[0]; P.sseg[8 * s + 7] = sh->bound;
  }
  w_sync();
  for (int b = lane; b < P.nStatBodies; b += 32) {
    const KShapeBody* sb = P.sbody + b;
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int k = 0; k < sb->n_shapes; ++k) {
      const KShape* sh = P.sshape + sb->first_shape + k;
      if (!sh->active0) continue;
      double smn[3], smx[3];
      k_shape_aabb(sh->type, sh->ext, P.sshapeA + 12 * (sb->first_shape + k), smn, smx);
      for (int q = 0; q < 3; ++q) { if (smn[q] < mn[q]) mn[q] = smn[q]; if (smx[q] > mx[q]) mx[q] = smx[q]; }
    }
    for (int q = 0; q < 3; ++q) { P.sbodyAABB[6 * b + q] = mn[q]; P.sbodyAABB[6 * b + 3 + q] = mx[q]; }
  }
}
