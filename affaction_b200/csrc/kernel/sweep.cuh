// One candidate rollout per THREAD ("sweep" kernel): 32 rollouts per warp, every lane busy in
// every phase, no shuffles, ballots or barriers.  Same algorithm and the same floating-point
// operations, in the same order, as k_rollout() of rollout.cuh (one rollout per warp), which is
// the kernel for small batches: the two produce identical bits, and both are checked against the
// CPU oracle with `==`.
//
//   TrajectoryPredictor::predict     (reference src/TrajectoryPredictor.cpp:79-406)
//   TrajectoryPredictor::computeIK   (:573-753)
//   TrajectoryPredictor::checkState  (:756-886)
//
// Why a second mapping: a warp that carries ONE rollout has 7 tasks, 14 trajectory dimensions,
// 23 joints and 2 x 12 transform elements to spread over 32 lanes and pays a barrier or shuffle
// at every hand-over (9.4 k warp instructions per step at 17.6 active lanes, round 1).  A sweep of
// 1e4 .. 1e6 perturbed candidates has no shortage of independent work: with one rollout per lane
// the loops over joints / nodes / pairs are warp-uniform (the plan is shared), only the data
// differ, and the per-rollout instruction count drops by an order of magnitude.
//
// State: per-thread arrays below (local memory: interleaved by lane, so a warp's access to
// element i of 32 rollouts is one coalesced 256-byte transaction, cached in L1 / L2).  The
// host plan compiler orders the candidates of a batch by structure (task slice, activation
// and constraint schedule) so that the lanes of a warp take the same branches.
//
// This file is included after rollout.cuh inside the same namespace and uses its helpers.
//
// AFFT_PHASE: the large phases (FK, collision model, task geometry) have two call sites each (initial
// state, step loop).  Inline by default; -DAFFT_PHASE_NOINLINE compiles one copy of each (smaller code,
// slower: pointers to the per-thread arrays turn into generic addresses).
// T_SYNC(): block-wide barrier between the phases of a step.  The warps of a block then walk the step
// loop within one phase of each other and share the instruction-cache lines of that phase (the loop body
// is far larger than the instruction caches; warps left to drift apart starve on instruction fetch).
// Every thread of the block reaches every barrier: threads without a candidate, or whose candidate is
// finished, skip the work between barriers but not the barriers.
#ifndef T_SYNC
#if defined(AFF_EMULATE) || !defined(AFFT_BLOCK_SYNC)
#define T_SYNC() do { } while (0)
#define T_BLOCK_MAX(v) (v)
#define T_BLOCK_ANY(p) (p)
#else
#define T_SYNC() __syncthreads()
#define T_BLOCK_MAX(v) t_block_max(v)
#define T_BLOCK_ANY(p) __syncthreads_or(p)
AFF_DEV int t_block_max(int v)
{
  __shared__ int smax;
  if (threadIdx.x == 0) smax = 0;
  __syncthreads();
  atomicMax(&smax, v);
  __syncthreads();
  return smax;
}
#endif
#endif

// -DAFFT_PROFILE (tuning builds only): thread 0 of block 0 adds the clock cycles between the phase barriers
#if defined(AFFT_PROFILE) && !defined(AFF_EMULATE)
__device__ unsigned long long g_phaseClocks[16];
#define T_MARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = clock64(); g_phaseClocks[i] += (unsigned long long)(now_ - tmark_); tmark_ = now_; } } while (0)
#define T_MARK_INIT long long tmark_ = clock64()
#define T_SUBMARK(i) do { if (blockIdx.x == 0 && threadIdx.x == 0) { const long long now_ = clock64(); g_phaseClocks[i] += (unsigned long long)(now_ - smark_); smark_ = now_; } } while (0)
#define T_SUBMARK_INIT long long smark_ = clock64()
#else
#define T_MARK(i) do { } while (0)
#define T_MARK_INIT do { } while (0)
#define T_SUBMARK(i) do { } while (0)
#define T_SUBMARK_INIT do { } while (0)
#endif

// barrier k of the step loop (0: loop head, 1: A|B, 2: B|C, 3: C|D, 4: D|E, 5: E|F); AFFT_SYNC_MASK selects which
// are real (tuning builds; default: all)
#ifndef AFFT_SYNC_MASK
#define AFFT_SYNC_MASK 0x15      /* measured round 2: loop head, B|C and D|E are enough to keep the warps of a block within a phase of each other */
#endif
#define T_SYNCI(k) do { if ((AFFT_SYNC_MASK >> (k)) & 1) T_SYNC(); } while (0)

#ifndef AFFT_PHASE
#ifdef AFFT_PHASE_NOINLINE
#define AFFT_PHASE AFF_NOINLINE
#else
#define AFFT_PHASE AFF_DEV
#endif
#endif

struct TState
{
  double nodes[12 * AFFT_MAX_SLOTS];
  unsigned chain[AFFT_MAX_SLOTS];
  int tdesc[8 * AFFT_MAX_TASKS];
  double q[AFFT_MAX_NJ], qdot[AFFT_MAX_NJ], vbuf[AFFT_MAX_NJ], ex[AFFT_MAX_NJ], dqv[AFFT_MAX_NJ];
  unsigned nz[AFFT_MAX_NJ];
  double st[3 * AFFT_MAX_XDIM], xdes[AFFT_MAX_XDIM], xcur[AFFT_MAX_XDIM], dxv[AFFT_MAX_XDIM];
  int viaHead[AFFT_MAX_XDIM], viaEnd[AFFT_MAX_XDIM], taskOfDim[AFFT_MAX_XDIM];
  int actHead[AFFT_MAX_TASKS], actBegin[AFFT_MAX_TASKS], actEnd[AFFT_MAX_TASKS], rowOf[AFFT_MAX_TASKS];
  double geo[AFFK_GEO * AFFT_MAX_TASKS];
  double J[AFFT_MAX_M * AFFT_MAX_NJ];
  double L[AFFT_MAX_M * AFFT_MAX_M], rhs[AFFT_MAX_M], linv[AFFT_MAX_M];
  unsigned rmask[AFFT_MAX_M];
  double objrel[12 * AFFK_MAX_OBJ], objq6[6 * AFFK_MAX_OBJ];
  double shp[12 * AFFT_MAX_DSHAPE], baabb[6 * AFFT_MAX_DBODY], taabb[6 * AFFT_MAX_TREES];
  int btree[AFFT_MAX_DBODY];
  int live[2 * AFFK_MAX_NEAR];
  int ckey[AFFT_MAX_COSTLY]; double cdist[AFFT_MAX_COSTLY];
  double via[AFFK_MAX_VIA_UNK * (AFFK_MAX_VIA_UNK + 1) + AFFK_MAX_VIA_UNK + 1];
  double tsum[16];
};

// Context of the shared helpers (k_slot, k_jac_cols, k_task_geo, k_task_J, k_pair_live ...).
struct TCtx
{
  const KPlan* P;
  TState* S;
  double* nodes;
  unsigned* chain;
  int objParentSlot0, objParentSlot1;
  int objParentDynamic0, objParentDynamic1;
  unsigned objChain0, objChain1;
  int objTree0, objTree1;
  int objMode0, objMode1;
  unsigned objHasRel;
  AFF_MEMBER const double* qv() const { return S->q; }
  AFF_MEMBER const double* shpv() const { return S->shp; }
};

// T o P with the operation order of k_fk(): every element is one fma chain seeded with the
// parent's origin element (translation) or 0.0 (rotation).
AFF_DEV void t_compose(double* out, const double* T, const double* Pp)
{
  double r[12];
  for (int e = 0; e < 3; ++e) r[e] = fma(T[2], Pp[9 + e], fma(T[1], Pp[6 + e], fma(T[0], Pp[3 + e], Pp[e])));
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j)
      r[3 + 3 * i + j] = fma(T[3 + 3 * i + 2], Pp[9 + j], fma(T[3 + 3 * i + 1], Pp[6 + j], fma(T[3 + 3 * i], Pp[3 + j], 0.0)));
  for (int e = 0; e < 12; ++e) out[e] = r[e];
}

// World record of dynamic shape s from the frame An of its body (the arithmetic of k_collision step 1).
AFF_DEV void t_shape_record(const KPlan& P, int s, const double* An, double* W)
{
  const KShape* sh = P.dshape + s;
  const int type = ldg(&sh->type);
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

// RcsGraph_setState for the dynamic nodes, parents first: nodes that do not hang below a
// re-parentable object in index order, then the objects and their subtrees (an object's current
// parent may be any node of the first group).
AFFT_PHASE void t_fk(TCtx& c, bool init)
{
  const KPlan& P = *c.P;
  TState& S = *c.S;
  // A holds the frame of node `held` (the node computed last): along a kinematic chain the parent of the next
  // node is that node, and its frame is taken from registers instead of being read back from the node array
  double A[12];
  int held = -1;
  for (int pass = 0; pass < 2; ++pass) {
    if (pass == 1 && P.nObj == 0) break;
    for (int n = 0; n < P.nDyn; ++n) {
      const int uo = ldg(&P.dyn_under[n]);
      if ((uo != 0) != (pass == 1)) continue;
      const KNode* nd = P.dyn + n;
      const int kind = ldg(&nd->kind);
      double* out = S.nodes + 12 * n;
      // the world records of the shapes that hang on this node are made here, from the frame in registers
      // (k_collision step 1), instead of re-reading the node array in the collision phase
      const int sh0 = ldg(&P.node_shape_off[n]), sh1 = ldg(&P.node_shape_off[n + 1]);
      if (kind == AFFK_NODE_OBJECT) {
        const int slot = ldg(&nd->qidx);
        const int ps = OBJ_GET(c, objParentSlot, slot);
        if (((c.objHasRel >> slot) & 1u) || init || OBJ_GET(c, objParentDynamic, slot)) {
          if (ps != held) { const double* Pp = S.nodes + 12 * ps; for (int e = 0; e < 12; ++e) A[e] = Pp[e]; }
          if ((c.objHasRel >> slot) & 1u) t_compose(A, S.objrel + 12 * slot, A);
          else for (int k = 0; k < 6; ++k) k_apply_joint(A, k, S.objq6[6 * slot + k]);
        } else {
          for (int e = 0; e < 12; ++e) A[e] = out[e];        // static parent, not re-parented: the pose stands
        }
      } else {
        // parent: a dynamic node, or one of the static frames copied behind the dynamic nodes
        const int ps = ldg(&P.dyn_pslot[n]);
        if (ps != held) { const double* Pp = S.nodes + 12 * ps; for (int e = 0; e < 12; ++e) A[e] = Pp[e]; }
        if (!(kind == AFFK_NODE_FIXED && ldg(&nd->identityT))) {
          double T[12];
          for (int e = 0; e < 12; ++e) T[e] = ldg(&nd->T[e]);
          t_compose(A, T, A);
          if (kind != AFFK_NODE_FIXED) {
            const int jt = ldg(&nd->jtype);
            double s, co;
            if (kind == AFFK_NODE_JOINT_IK) {
              const double qv = S.q[ldg(&nd->qidx)];
              if (jt >= AFF_JNT_ROT_X) aff_sincos(qv, &s, &co); else { s = qv; co = 0.0; }
            } else {
              s = (jt >= AFF_JNT_ROT_X) ? ldg(&P.dynSC[2 * n]) : ldg(&P.dynQ[n]);
              co = ldg(&P.dynSC[2 * n + 1]);
            }
            if (jt >= AFF_JNT_ROT_X) {
              const int d = jt - AFF_JNT_ROT_X, a = (d + 1) % 3, b = (d + 2) % 3;
              for (int k = 0; k < 3; ++k) {
                const double ra = A[3 + 3 * a + k], rb = A[3 + 3 * b + k];
                A[3 + 3 * a + k] = fma(s, rb, co * ra);
                A[3 + 3 * b + k] = fma(co, rb, -(s * ra));
              }
            } else {
              for (int k = 0; k < 3; ++k) A[k] = fma(s, A[3 + 3 * jt + k], A[k]);
            }
          }
        }
      }
      held = n;
      if (ldg(&P.dyn_store[n])) for (int e = 0; e < 12; ++e) out[e] = A[e];      // frames nobody reads again stay in registers
      for (int k = sh0; k < sh1; ++k) { const int si = ldg(&P.node_shape_idx[k]); t_shape_record(P, si, A, S.shp + 12 * si); }
    }
  }
}

// Rare path: a body pair found by the per-step AABB test (out of line: it carries every distance function).
AFF_NOINLINE double t_pair_distance(const TCtx& c, int ia, int ib) { return k_pair_distance(c, ia, ib); }

// ControllerBase::computeCollisionModel (k_collision of rollout.cuh, one thread).
AFFT_PHASE void t_collision(TCtx& c, double need, KCollResult* out)
{
  const KPlan& P = *c.P;
  TState& S = *c.S;
  const double thr = P.bp_threshold;
  double* shp = S.shp; double* baabb = S.baabb; double* taabb = S.taabb;
  int* btree = S.btree;
  T_SUBMARK_INIT;
  // 1. the world records of the dynamic shapes were written by the FK walk (t_fk)
  T_SUBMARK(8);
  // 2. per dynamic body: active flag, tree, AABB; 3. the tree boxes are the running min / max over their bodies
  for (int i = 0; i < 6 * P.nTrees; ++i) taabb[i] = (i % 6 < 3) ? DBL_MAX : -DBL_MAX;
  for (int bi = 0; bi < P.nDynBodies; ++bi) {
    const KShapeBody* b = P.dbody + bi;
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
    for (int q = 0; q < 3; ++q) { baabb[6 * bi + q] = mn[q]; baabb[6 * bi + 3 + q] = mx[q]; }
    const int bt = act ? (os ? OBJ_GET(c, objTree, os - 1) : ldg(&b->tree)) : -2;
    btree[bi] = bt;
    if (bt >= 0)
      for (int q = 0; q < 3; ++q) { if (mn[q] < taabb[6 * bt + q]) taabb[6 * bt + q] = mn[q]; if (mx[q] > taabb[6 * bt + 3 + q]) taabb[6 * bt + 3 + q] = mx[q]; }
  }
  T_SUBMARK(9);
  T_SUBMARK(10);
  // 4. static bodies that are not explicit: tree box first, then the tree's bodies
  int nLive = 0;
  double uni[6] = {DBL_MAX, DBL_MAX, DBL_MAX, -DBL_MAX, -DBL_MAX, -DBL_MAX};     // union of the tree boxes: a pure pre-filter
  for (int t = 0; t < P.nTrees; ++t)
    for (int k = 0; k < 3; ++k) { if (taabb[6 * t + k] < uni[k]) uni[k] = taabb[6 * t + k]; if (taabb[6 * t + 3 + k] > uni[3 + k]) uni[3 + k] = taabb[6 * t + 3 + k]; }
  for (int i = 0; i < P.nStatNear; ++i) {
    double bb[6];
    for (int q = 0; q < 6; ++q) bb[q] = ldg(&P.sbodyAABB[6 * i + q]);
    if (!k_aabb_overlap(uni, bb, thr)) continue;
    int near = 0;
    for (int t = 0; t < P.nTrees; ++t) if (k_aabb_overlap(taabb + 6 * t, bb, thr)) near = 1;
    if (!near) continue;
    for (int b = 0; b < P.nDynBodies; ++b) {
      if (btree[b] < 0) continue;
      if (!k_aabb_overlap(baabb + 6 * b, bb, thr)) continue;
      if (nLive < AFFK_MAX_NEAR) { S.live[2 * nLive] = b; S.live[2 * nLive + 1] = -1 - i; }
      ++nLive;
    }
  }
  const int nearOverflow = nLive > AFFK_MAX_NEAR;
  if (nLive > AFFK_MAX_NEAR) nLive = AFFK_MAX_NEAR;
  T_SUBMARK(11);
  // 5. narrow phase
  double best = DBL_MAX; int bestKey = 0x7fffffff; int nCostly = 0;
  const double cullMargin = 1.0e-9;
  const double cullNeed = (need > thr) ? need : thr;
#define T_ACCOUNT(d, key)                                                                                 \
  do {                                                                                                    \
    if ((d) < best || ((d) == best && (key) < bestKey)) { best = (d); bestKey = (key); }                  \
    if (thr > 0.0 && (d) < thr) { if (nCostly < AFFT_MAX_COSTLY) { S.ckey[nCostly] = (key); S.cdist[nCostly] = (d); } ++nCostly; } \
  } while (0)
  // Tree-level pre-filter: two broad-phase trees whose boxes are at least cullNeed apart cannot hold a pair
  // that becomes the minimum or carries a cost (a tree's box contains every active shape of its bodies), so
  // their pairs are skipped before any shape record is read - in a `get`, all of arm x arm.  A pure filter:
  // pairs it drops would have been dropped by the bounding-sphere test or lost against `need` anyway.
  unsigned long long treeFar = 0ull;
  {
    const double lim = cullNeed + cullMargin;
    for (int ta = 0; ta < P.nTrees; ++ta)
      for (int tb = ta + 1; tb < P.nTrees; ++tb) {
        double g2 = 0.0;
        for (int k = 0; k < 3; ++k) {
          const double g1 = taabb[6 * ta + k] - taabb[6 * tb + 3 + k], g2k = taabb[6 * tb + k] - taabb[6 * ta + 3 + k];
          const double g = g1 > g2k ? g1 : g2k;
          if (g > 0.0) g2 += g * g;
        }
        if (g2 >= lim * lim) treeFar |= (1ull << (ta * 8 + tb)) | (1ull << (tb * 8 + ta));
      }
  }
#define T_TREES_FAR(meta) (!KP_RULE(meta) && KP_TREE_A(meta) >= 0 && KP_TREE_B(meta) >= 0 && ((treeFar >> (KP_TREE_A(meta) * 8 + KP_TREE_B(meta))) & 1ull))
  int lastA = -1;
  double A[8];
  for (int i = 0; i < P.nSS; ++i) {
    const KPair pr = ldg_pair(&P.pairsSS[i]);
    if (T_TREES_FAR(pr.meta)) continue;
    if (!k_pair_live(c, pr.meta, baabb, thr)) continue;
    if (pr.a != lastA) { const double* Ap = shp + 12 * pr.a; for (int k = 0; k < 8; ++k) A[k] = Ap[k]; lastA = pr.a; }
    const double rA = A[6];
    double pB[3], dB[3], rB, bB;
    if (pr.b >= 0) { const double* B = shp + 12 * pr.b; for (int k = 0; k < 3; ++k) { pB[k] = B[k]; dB[k] = B[3 + k]; } rB = B[6]; bB = B[7]; }
    else { const double* B = P.sseg + 8 * (-1 - pr.b); for (int k = 0; k < 3; ++k) { pB[k] = ldg(&B[k]); dB[k] = ldg(&B[3 + k]); } rB = ldg(&B[6]); bB = ldg(&B[7]); }
    double cd2 = 0.0;
    for (int k = 0; k < 3; ++k) { const double dc = (A[k] + 0.5 * A[3 + k]) - (pB[k] + 0.5 * dB[k]); cd2 += dc * dc; }
    const double lim = cullNeed + (A[7] + bB) + cullMargin;
    if (cd2 >= lim * lim) continue;
    const double d = KP_SWAP(pr.meta) ? (k_seg_seg(pB, dB, A, A + 3) - rB) - rA : (k_seg_seg(A, A + 3, pB, dB) - rA) - rB;
    T_ACCOUNT(d, pr.key);
  }
  T_SUBMARK(12);
  for (int i = 0; i < P.nSB; ++i) {
    const KPair pr = ldg_pair(&P.pairsSB[i]);
    if (T_TREES_FAR(pr.meta)) continue;
    if (!k_pair_live(c, pr.meta, baabb, thr)) continue;
    double pS[3], dS[3], rS, bS, cB[3], bB;
    if (pr.a >= 0) {
      if (pr.a != lastA) { const double* Ap = shp + 12 * pr.a; for (int k = 0; k < 8; ++k) A[k] = Ap[k]; lastA = pr.a; }
      for (int k = 0; k < 3; ++k) { pS[k] = A[k]; dS[k] = A[3 + k]; }
      rS = A[6]; bS = A[7];
    }
    else { const double* As = P.sseg + 8 * (-1 - pr.a); for (int k = 0; k < 3; ++k) { pS[k] = ldg(&As[k]); dS[k] = ldg(&As[3 + k]); } rS = ldg(&As[6]); bS = ldg(&As[7]); }
    if (pr.b >= 0) { const double* B = shp + 12 * pr.b; for (int k = 0; k < 3; ++k) cB[k] = B[k]; bB = ldg(&P.dshape[pr.b].bound); }
    else { const double* B = P.sseg + 8 * (-1 - pr.b); for (int k = 0; k < 3; ++k) cB[k] = ldg(&B[k]); bB = ldg(&B[7]); }
    double cd2 = 0.0;
    for (int k = 0; k < 3; ++k) { const double dc = (pS[k] + 0.5 * dS[k]) - cB[k]; cd2 += dc * dc; }
    const double lim = cullNeed + (bS + bB) + cullMargin;
    if (cd2 >= lim * lim) continue;
    const bool dynB = pr.b >= 0;
    const double* Bs = shp + 12 * (dynB ? pr.b : 0);
    const double* Bg = P.sshapeA + 12 * (dynB ? 0 : -1 - pr.b);
    const KShape* sb = dynB ? P.dshape + pr.b : P.sshape + (-1 - pr.b);
    double rel[3], pl[3], dl[3];
    for (int k = 0; k < 3; ++k) rel[k] = pS[k] - (dynB ? Bs[k] : ldg(&Bg[k]));
    for (int r = 0; r < 3; ++r) {
      const double r0 = dynB ? Bs[3 + 3 * r] : ldg(&Bg[3 + 3 * r]), r1 = dynB ? Bs[4 + 3 * r] : ldg(&Bg[4 + 3 * r]),
                   r2 = dynB ? Bs[5 + 3 * r] : ldg(&Bg[5 + 3 * r]);
      pl[r] = fma(r2, rel[2], fma(r1, rel[1], r0 * rel[0]));
      dl[r] = fma(r2, dS[2], fma(r1, dS[1], r0 * dS[0]));
    }
    const int typeB = ldg(&sb->type);
    double ext[3];
    for (int k = 0; k < 3; ++k) ext[k] = ldg(&sb->ext[k]);
    double d;
    if (typeB == AFF_SHAPE_BOX) { const double h[3] = {0.5 * ext[0], 0.5 * ext[1], 0.5 * ext[2]}; d = k_seg_box(pl, dl, h) - rS; }
    else d = k_seg_cyl(pl, dl, 0.5 * ext[2], ext[0]) - rS;
    T_ACCOUNT(d, pr.key);
  }
  T_SUBMARK(13);
  for (int i = 0; i < nLive; ++i) {
    const int ia = S.live[2 * i], ib = S.live[2 * i + 1];
    const double d = t_pair_distance(c, ia, ib);
    const int bodyA = ldg(&P.dbody[ia].body);
    const int bodyB = (ib >= 0) ? ldg(&P.dbody[ib].body) : ldg(&P.sbody[-1 - ib].body);
    const int b1 = bodyA < bodyB ? bodyA : bodyB, b2 = bodyA < bodyB ? bodyB : bodyA;
    const int key = (b1 << 15) | b2;
    T_ACCOUNT(d, key);
  }
#undef T_ACCOUNT
#undef T_TREES_FAR
  KCollResult res;
  if (best < need) {
    res.minDist = best;
    if (!(best < DBL_MAX)) { res.minB1 = -1; res.minB2 = -1; }
    else { res.minB1 = bestKey >> 15; res.minB2 = bestKey & 0x7fff; }
  } else { res.minDist = DBL_MAX; res.minB1 = -1; res.minB2 = -1; }
  // cost: one term per body pair (its closest shape pair), in ascending (b1, b2) order
  res.cost = 0.0;
  res.overflow = (nearOverflow || nCostly > AFFT_MAX_COSTLY) ? 1 : 0;
  if (nCostly > 0) {
    const int n = nCostly < AFFT_MAX_COSTLY ? nCostly : AFFT_MAX_COSTLY;
    for (;;) {
      int kmin = 0x7fffffff;
      for (int i = 0; i < n; ++i) if (S.ckey[i] < kmin) kmin = S.ckey[i];
      if (kmin == 0x7fffffff) break;
      double d = DBL_MAX;
      for (int i = 0; i < n; ++i) if (S.ckey[i] == kmin) { if (S.cdist[i] < d) d = S.cdist[i]; S.ckey[i] = 0x7fffffff; }
      res.cost = res.cost + ((d < 0.0) ? 1.0 - (2.0 * d) / thr : ((d - thr) / thr) * ((d - thr) / thr));
    }
  }
  *out = res;
}

// Task geometry of every task / task error of one task, out of line (two call sites each).
AFFT_PHASE void t_task_geo_all(const TCtx& c, int nTasks, const KTask* tasks)
{
  TState& S = *c.S;
  for (int t = 0; t < nTasks; ++t) k_task_geo(c, S.tdesc + 8 * t, tasks + t, S.geo + AFFK_GEO * t, S.xcur);
}
AFFT_PHASE void t_task_dx(const TCtx& c, int t, const KTask* tasks, double* dx)
{
  TState& S = *c.S;
  k_task_dx(S.tdesc + 8 * t, tasks + t, S.geo + AFFK_GEO * t, S.xcur, S.xdes, dx);
}

// The rollout of candidate `cand` on the calling thread.
AFF_DEV void t_rollout(const KPlan& P, int cand, TState& S, bool alive)
{
  TCtx c;
  c.P = &P; c.S = &S; c.nodes = S.nodes; c.chain = S.chain;
  const KCand& C = P.cands[cand];
  const int nJ = P.nJ;
  const double dt = P.opts.dt;
  const int nTasks = ldg(&C.n_tasks), xdim = ldg(&C.xdim);
  const KTask* tasks = P.tasks + ldg(&C.first_task);
  const int nSlots = P.nDyn + P.nStatPar;
  const int JS = nJ, LS = AFFT_MAX_M;
  double* q = S.q; double* qdot = S.qdot; double* J = S.J; double* L = S.L; double* rhs = S.rhs;
  double* dxv = S.dxv; double* st = S.st; double* xdes = S.xdes; double* xcur = S.xcur; double* geo = S.geo;

  for (int j = 0; j < nJ; ++j) { q[j] = ldg(&P.jq_init[j]); qdot[j] = 0.0; }
  for (int t = 0; t < nTasks; ++t) {
    const KTask* tk = tasks + t;
    int* td = S.tdesc + 8 * t;
    const int kind = ldg(&tk->kind);
    td[TD_KIND] = kind; td[TD_EFF] = ldg(&tk->eff); td[TD_REF] = ldg(&tk->ref); td[TD_FRAME] = ldg(&tk->frame);
    td[TD_AXIS] = ldg(&tk->axis); td[TD_XOFF] = ldg(&tk->xoff); td[TD_NJ] = ldg(&tk->n_joints);
    td[TD_DIM] = (kind == AFF_TASK_XYZ) ? 3 : ((kind == AFF_TASK_POLAR) ? 2 : ((kind == AFF_TASK_JOINTS) ? ldg(&tk->n_joints) : 1));
    S.actBegin[t] = ldg(&C.first_act) + ldg(&C.act_off[t]);
    S.actEnd[t] = ldg(&C.first_act) + ldg(&C.act_off[t + 1]);
    S.actHead[t] = S.actBegin[t];
  }
  for (int d = 0; d < xdim; ++d) {
    int mt = -1;
    for (int t = 0; t < nTasks; ++t) if (d >= S.tdesc[8 * t + TD_XOFF]) mt = t;
    S.taskOfDim[d] = mt;
    S.viaHead[d] = ldg(&C.first_via) + ldg(&C.via_off[d]);
    S.viaEnd[d] = ldg(&C.first_via) + ldg(&C.via_off[d + 1]);
    st[3 * d] = 0.0; st[3 * d + 1] = 0.0; st[3 * d + 2] = 0.0; xdes[d] = 0.0;
  }
  const KVia* vias = P.vias;
  const KAct* acts = P.acts;
  const KGraphCon* gcons = P.gcons + ldg(&C.first_gcon);
  const int nGcon = ldg(&C.n_gcon);
  unsigned gfired = 0u;

  // slots: static frames the loop refers to, chains
  for (int k = 0; k < P.nStatPar; ++k) {
    const int ref = ldg(&P.statpar_ref[k]);
    const double* src = (ref == AFFK_WORLD) ? P.statA + 12 * P.nStatic : P.statA + 12 * AFFK_STATIC_IDX(ref);
    for (int e = 0; e < 12; ++e) S.nodes[12 * (P.nDyn + k) + e] = ldg(&src[e]);
  }
  for (int i = 0; i < nSlots; ++i) S.chain[i] = (i < P.nDyn) ? ldg(&P.dyn_chain[i]) : 0u;
  // objects
  c.objHasRel = 0u;
  c.objParentSlot0 = c.objParentSlot1 = 0; c.objParentDynamic0 = c.objParentDynamic1 = 0; c.objChain0 = c.objChain1 = 0u;
  c.objTree0 = c.objTree1 = -1; c.objMode0 = c.objMode1 = 0;
  for (int s = 0; s < P.nObj; ++s) {
    const int ps = ldg(&P.obj_parent_slot[s]);
    OBJ_SET(c, objParentSlot, s, ps);
    OBJ_SET(c, objParentDynamic, s, ps < P.nDyn);
    OBJ_SET(c, objChain, s, S.chain[ps]);
    OBJ_SET(c, objTree, s, ldg(&P.obj_parent_tree[s]));
  }
  for (int i = 0; i < P.nDyn; ++i) { const int uo = ldg(&P.dyn_under[i]); if (uo) S.chain[i] = ldg(&P.dyn_chain[i]) | OBJ_GET(c, objChain, uo - 1); }
  for (int s = 0; s < P.nObj; ++s)
    for (int k = 0; k < 6; ++k) S.objq6[6 * s + k] = (ldg(&C.pose_slot) == s) ? ldg(&C.pose_q6[k]) : ldg(&P.obj_q6[6 * s + k]);

  // initial state
  t_fk(c, true);
  const int nSteps = ldg(&C.n_steps);
  const int qrows = P.qlog_rows;
  double* qlog = P.qlog ? P.qlog + (size_t)cand * qrows * nJ : nullptr;
  if (!alive) qlog = nullptr;
  if (qlog && qrows > 0) for (int j = 0; j < nJ; ++j) qlog[j] = q[j];

  int r_code = 0, r_first_code = 0, r_first_fail = -1, r_fail_step = -1, r_id0 = -1, r_id1 = -1, r_b1 = -1, r_b2 = -1;
  double r_minDist = DBL_MAX;
  KCollResult cm;
  t_collision(c, DBL_MAX, &cm);
  r_minDist = cm.minDist; r_b1 = cm.minB1; r_b2 = cm.minB2;
  unsigned r_flags = cm.overflow ? AFF_FLAG_LIST_OVERFLOW : 0u;

  // check(): checkLimits(jl, coll, speed, 0, 0, 0.01)  (src/TrajectoryPredictor.cpp:426-441)
  {
    bool ok0 = true;
    for (int j = 0; j < nJ; ++j) if (q[j] < ldg(&P.jq_min[j]) || q[j] > ldg(&P.jq_max[j])) ok0 = false;
    if (cm.minB1 >= 0 && cm.minDist < 0.01) ok0 = false;
    if (!ok0) {
      aff_result res;
      res.idx = cand; res.success = 0; res.code = AFF_CODE_INITIAL_STATE; res.first_code = AFF_CODE_INITIAL_STATE;
      res.first_fail_step = 0; res.fail_step = 0; res.fail_id0 = -1; res.fail_id1 = -1;
      res.min_dist_b1 = r_b1; res.min_dist_b2 = r_b2; res.n_steps = 0; res.reserved = r_flags; res.jmask_bits = 0ull;
      res.min_dist = r_minDist; res.jl_cost = 0.0; res.coll_cost = 0.0; res.track_err = 0.0;
      if (alive) P.results[cand] = res;
      alive = false;
    }
  }
  t_task_geo_all(c, nTasks, tasks);

  unsigned activeMask = 0u, prevMask = 0u, jmaskBits = 0u;
  unsigned emask = ~(P.rbj_static_chain);
  emask &= ~(c.objChain0 | c.objChain1);
  double tnow = 0.0, qFilt = 0.0, err = 0.0, jlSum = 0.0, collSum = 0.0, endTime = ldg(&C.end_abs);
  const double endAbs = ldg(&C.end_abs);
  const double motionDuration = endAbs - ldg(&C.start_abs);
  const double alpha = P.opts.alpha, lambda = P.opts.lambda, qFiltDecay = P.opts.q_filt_decay;
  int success = 1, rows = 1;

  const int loopSteps = T_BLOCK_MAX(alive ? nSteps : 0);
  double blending = 1.0, phaseScale = 0.0, alphaEff = 0.0;
  int m = 0, singular = 0, ikRes = 0, id0 = -1, id1 = -1, fast = 0;
  T_MARK_INIT;
  bool running = alive;      // early exit (opt-in) stops a failed rollout; its thread keeps meeting the block's barriers
  for (int iter = 0; iter < loopSteps; ++iter) {
    if (P.opts.early_exit && !T_BLOCK_ANY(running && iter < nSteps)) break;
    const bool act = running && iter < nSteps;
    const int successPrev = success;
    T_SYNCI(0);
    T_MARK(6);
    if (act) {   // ======== phase A: trajectories, constraints, activations, task errors
    // ---- task switch / qFilt (:166-180)
    if (prevMask != activeMask) qFilt = 1.0;
    if (qFiltDecay <= 0.0) qFilt = 0.0;
    else qFilt = k_clip(qFilt - (1.0 / qFiltDecay) * dt, 0.0, 1.0);

    // ---- tc->step(dt): trajectories
    const double tnext = tnow + dt;
    for (int d = 0; d < xdim; ++d) {
      const int mt = S.taskOfDim[d];
      const bool active = (activeMask >> mt) & 1u;
      bool switchesOn = false;
      for (int k = S.actHead[mt]; k < S.actEnd[mt]; ++k) { if (!(ldg(&acts[k].t) - tnext < AFFK_TRAJ_EPS)) break; switchesOn = ldg(&acts[k].on) != 0; }
      if (active || switchesOn) {
        double s3[3];
        if (!active) { s3[0] = xcur[d]; s3[1] = 0.0; s3[2] = 0.0; }
        else { s3[0] = st[3 * d]; s3[1] = st[3 * d + 1]; s3[2] = st[3 * d + 2]; }
        k_via_step(s3, vias + S.viaHead[d], S.viaEnd[d] - S.viaHead[d], tnow, dt, S.via, P.via_unk);
        st[3 * d] = s3[0]; st[3 * d + 1] = s3[1]; st[3 * d + 2] = s3[2];
        xdes[d] = s3[0];
      }
    }
    tnow = tnext;
    // graph constraints
    for (int k = 0; k < nGcon; ++k) {
      if ((gfired >> k) & 1u) continue;
      const double tc = ldg(&gcons[k].t);
      const int kind = ldg(&gcons[k].kind), slot = ldg(&gcons[k].slot);
      if (kind == AFF_CON_CONNECT) {
        if (tc - tnow < AFFK_TRAJ_EPS) {
          const int ps = ldg(&gcons[k].parent_slot);
          const int on = ldg(&P.obj_node[slot]);
          const double* Cc = k_slot(c, on); const double* Pp = k_slot(c, ps);
          double rel[12];
          for (int l = 0; l < 12; ++l) {
            double v;
            if (ldg(&gcons[k].on)) v = (l == 3 || l == 7 || l == 11) ? 1.0 : 0.0;
            else if (l < 3) {
              const double d0 = Cc[0] - Pp[0], d1 = Cc[1] - Pp[1], d2 = Cc[2] - Pp[2];
              v = fma(Pp[3 + 3 * l + 2], d2, fma(Pp[3 + 3 * l + 1], d1, Pp[3 + 3 * l] * d0));
            } else {
              const int i = (l - 3) / 3, j = (l - 3) % 3;
              v = fma(Cc[3 + 3 * i + 2], Pp[3 + 3 * j + 2], fma(Cc[3 + 3 * i + 1], Pp[3 + 3 * j + 1], Cc[3 + 3 * i] * Pp[3 + 3 * j]));
            }
            rel[l] = v;
          }
          for (int l = 0; l < 12; ++l) S.objrel[12 * slot + l] = rel[l];
          c.objHasRel |= 1u << slot;
          OBJ_SET(c, objParentSlot, slot, ps);
          OBJ_SET(c, objParentDynamic, slot, ps < P.nDyn);
          { const unsigned ch = S.chain[ps]; OBJ_SET(c, objChain, slot, ch); }
          const int pt = ldg(&gcons[k].parent_tree);
          { const int tr = (pt <= -2) ? OBJ_GET(c, objTree, -2 - pt) : pt; OBJ_SET(c, objTree, slot, tr); }
          for (int i = 0; i < P.nDyn; ++i) if (ldg(&P.dyn_under[i]) == slot + 1) S.chain[i] = ldg(&P.dyn_chain[i]) | OBJ_GET(c, objChain, slot);
          emask = ~(P.rbj_static_chain);
          emask &= ~(c.objChain0 | c.objChain1);
          gfired |= 1u << k;
        }
      } else {
        if (tc - tnow < 0.0) { const int md = ldg(&gcons[k].on) ? 1 : 2; OBJ_SET(c, objMode, slot, md); gfired |= 1u << k; }
      }
    }
    // activation points, via heads
    prevMask = activeMask;
    {
      unsigned nm = 0u;
      for (int t = 0; t < nTasks; ++t) {
        int myActive = (activeMask >> t) & 1u;
        int ah = S.actHead[t];
        const int ae = S.actEnd[t];
        while (ah < ae && ldg(&acts[ah].t) - tnow < AFFK_TRAJ_EPS) { myActive = ldg(&acts[ah].on); ++ah; }
        S.actHead[t] = ah;
        if (myActive) nm |= 1u << t;
      }
      activeMask = nm;
    }
    for (int d = 0; d < xdim; ++d) { int vh = S.viaHead[d]; const int ve = S.viaEnd[d]; while (vh < ve && ldg(&vias[vh].t) - tnow < AFFK_TRAJ_EPS) ++vh; S.viaHead[d] = vh; }
    {
      const double rem = endAbs - tnow;
      endTime = rem > 0.0 ? rem : 0.0;
    }
    // blending
    blending = 1.0;
    for (int t = 0; t < nTasks; ++t) {
      double bl = 1.0;
      const int ah = S.actHead[t];
      if (ah < S.actEnd[t]) { const double h = ldg(&acts[ah].horizon); if (h > 0.0) bl = k_clip((ldg(&acts[ah].t) - tnow) / h, 0.0, 1.0); }
      if (ah > S.actBegin[t]) {
        const double lastSwitchT = ldg(&acts[ah - 1].t), lastSwitchH = ldg(&acts[ah - 1].horizon);
        if (lastSwitchH > 0.0) { const double up = k_clip((tnow - lastSwitchT) / lastSwitchH, 0.0, 1.0); if (up < bl) bl = up; }
      }
      if (bl < blending) blending = bl;
    }
    if (activeMask == 0u) blending = 0.0;
    const double phase = (motionDuration > 0.0) ? 1.0 - (endTime / motionDuration) : 0.0;
    phaseScale = aff_sin(AFF_PI * phase);
    alphaEff = blending * alpha;

    // ================= computeIK =================
    m = 0;
    for (int t = 0; t < nTasks; ++t) { S.rowOf[t] = m; if ((activeMask >> t) & 1u) m += S.tdesc[8 * t + TD_DIM]; }
    for (int t = 0; t < nTasks; ++t)
      if ((activeMask >> t) & 1u) t_task_dx(c, t, tasks, dxv + S.rowOf[t]);
    }
    T_SYNCI(1);
    T_MARK(0);
    if (act) {   // ======== phase B: Jacobian columns, null-space gradient
    // Elbow / wrist terms (:448-570): the factors do not depend on the joint, only the Jacobian entries do
    // (the warp-per-candidate kernel recomputes them on every lane; same operations, same values).
    double nsF[6]; int nsBody[6], nsRow[6], nN = 0;
    unsigned nsChain = 0u;
    if (P.ns_base >= 0) {
      const double* Ab = k_slot(c, P.ns_base);
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
          nsF[nN] = f; nsBody[nN] = el; nsRow[nN] = 1; ++nN;
          nsChain |= S.chain[el];
        }
      }
      const int nElbow = nN;
      for (int side = 0; side < 2; ++side) {
        const int wr = P.ns_hand[side];
        if (wr < 0) continue;
        const double* Aw = k_slot(c, wr);
        double dW[3];
        for (int k = 0; k < 3; ++k) dW[k] = Aw[k] - Ab[k];
        const double x = fma(Ab[3 + 2], dW[2], fma(Ab[3 + 1], dW[1], Ab[3] * dW[0]));
        const double z = fma(Ab[3 + 6 + 2], dW[2], fma(Ab[3 + 6 + 1], dW[1], Ab[3 + 6] * dW[0]));
        const double dBound = (z > 1.0) ? 0.4 * (z - 1.0) : 0.0;
        nsF[nN] = k_clip(x - (0.25 + dBound), -0.5, 0.0) * 5.0; nsBody[nN] = wr; nsRow[nN] = 0; ++nN;
        nsChain |= S.chain[wr];
      }
      nsChain |= S.chain[P.ns_base];
      (void)nElbow;
    }
    // joints that can have a non-zero entry in the rows of each active task (everything else is exactly zero,
    // is never read through the row masks, and is not computed)
    unsigned taskChain[AFFT_MAX_TASKS];
    for (int t = 0; t < nTasks; ++t) {
      unsigned tc = 0u;
      if ((activeMask >> t) & 1u) {
        const int* td = S.tdesc + 8 * t;
        if (td[TD_KIND] == AFF_TASK_JOINTS) { for (int i = 0; i < td[TD_NJ]; ++i) { const int ji = ldg(&tasks[t].jidx[i]); if (ji >= 0) tc |= 1u << ji; } }
        else {
          if (td[TD_EFF] >= 0) tc |= S.chain[td[TD_EFF]];
          if (td[TD_REF] >= 0) tc |= S.chain[td[TD_REF]];
          if (td[TD_FRAME] >= 0) tc |= S.chain[td[TD_FRAME]];
        }
      }
      taskChain[t] = tc;
    }
    // per joint: J column, dH (joint-limit gradient + elbow / wrist terms)
    unsigned aMask = 0u;
    unsigned anyChain = nsChain;
    for (int t = 0; t < nTasks; ++t) anyChain |= taskChain[t];
    for (int r = 0; r < m; ++r) S.rmask[r] = 0u;
    for (int j = 0; j < nJ; ++j) {
      KJointFrame jf;
      const int jtype = ldg(&P.j_type[j]);
      jf.valid = 1; jf.type = jtype;
      for (int k = 0; k < 3; ++k) { jf.ax[k] = 0.0; jf.org[k] = 0.0; }
      if ((anyChain >> j) & 1u) {       // the frame of a joint outside every chain is never used (and not kept: KPlan::dyn_store)
        const double* A = k_slot(c, ldg(&P.j_node[j]));
        const int ar = (jtype < AFF_JNT_ROT_X) ? jtype : jtype - AFF_JNT_ROT_X;
        for (int k = 0; k < 3; ++k) { jf.ax[k] = A[3 + 3 * ar + k]; jf.org[k] = A[k]; }
      }
      unsigned nz = 0u;
      for (int t = 0; t < nTasks; ++t) {
        if (!((taskChain[t] >> j) & 1u)) continue;
        const int row = S.rowOf[t];
        nz |= k_task_J(c, jf, j, S.tdesc + 8 * t, tasks + t, geo + AFFK_GEO * t, J + row * JS, JS) << row;
      }
      S.nz[j] = nz;
      if (nz) aMask |= 1u << j;
      { unsigned mk = nz; while (mk) { const int r = w_ffs(mk) - 1; mk &= mk - 1u; S.rmask[r] |= 1u << j; } }
      const double jq0 = ldg(&P.jq0[j]);
      const double dqc = q[j] - jq0;
      const double rr0 = (dqc > 0.0) ? ldg(&P.jq_max[j]) - jq0 : jq0 - ldg(&P.jq_min[j]);
      S.vbuf[j] = ((ldg(&P.jwjl[j]) * dqc) / (rr0 * rr0)) * alphaEff;     // dH, joint-limit part
      double ex = 0.0;
      if (P.ns_base >= 0) {
        // tmp accumulates (te - tr) * f per term in the order of the reference; a joint outside both chains
        // contributes (0 - 0) * f = +-0, which leaves the sum unchanged, and is skipped
        double tmp = 0.0, tmpW = 0.0;
        if ((nsChain >> j) & 1u) {
          const int base = P.ns_base;
          const double* Ab = k_slot(c, base);
          for (int k = 0; k < nN; ++k) {
            const int bd = nsBody[k];
            if (!(((S.chain[bd] | S.chain[base]) >> j) & 1u)) continue;
            double te[3], re[3], tr[3], rr[3];
            k_jac_cols(c, jf, j, bd, k_slot(c, bd), te, re);
            k_jac_cols(c, jf, j, base, Ab, tr, rr);
            if (nsRow[k]) tmp = tmp + (te[1] - tr[1]) * nsF[k];
            else tmpW = tmpW + (te[0] - tr[0]) * nsF[k];
          }
        }
        ex = ex + tmp;
        ex = (ex + tmpW) * alphaEff;
      }
      S.ex[j] = ex;
    }
    // RcsGraph_checkJointSpeeds(dH_extra, dt, IK)
    {
      double scale = 1.0;
      for (int j = 0; j < nJ; ++j) {
        const double js = ldg(&P.jspeed[j]);
        if (js < AFFK_NO_LIMIT) { const double a = fabs(S.ex[j]), mx = js * dt; if (a > mx) { const double sc = mx / a; if (sc < scale) scale = sc; } }
      }
      const bool scaled = scale < 1.0;
      for (int j = 0; j < nJ; ++j) {
        double ex = S.ex[j];
        if (scaled) ex = ex * (0.99999 * scale);
        double dHv = S.vbuf[j] + ex * phaseScale;
        if (!(((aMask | emask) >> j) & 1u)) dHv = dHv * 0.0;
        S.vbuf[j] = ldg(&P.jwmetric[j]) * dHv;          // v = W^-1 dH
      }
    }
    jmaskBits |= aMask;
    }
    T_SYNCI(2);
    T_MARK(1);
    if (act) {   // ======== phase C: right inverse, limits, integration
    // ---- solveRightInverse
    for (int i = 0; i < m; ++i) {
      double acc = dxv[i];
      unsigned mk = S.rmask[i];
      while (mk) { const int j = w_ffs(mk) - 1; mk &= mk - 1u; acc = fma(J[i * JS + j], S.vbuf[j], acc); }
      rhs[i] = acc;
    }
    for (int i = 0; i < m; ++i)
      for (int k = 0; k <= i; ++k) {
        double acc = 0.0;
        unsigned mk = S.rmask[i] & S.rmask[k];
        while (mk) { const int j = w_ffs(mk) - 1; mk &= mk - 1u; acc = fma(J[i * JS + j] * ldg(&P.jwmetric[j]), J[k * JS + j], acc); }
        if (i == k) acc = acc + lambda;
        L[i * LS + k] = acc;
      }
    // Cholesky with reciprocal pivots (column by column, as the one-lane-per-row version)
    singular = 0;
    for (int j = 0; j < m; ++j) {
      double tj = L[j * LS + j];
      for (int k = 0; k < j; ++k) tj = fma(-L[j * LS + k], L[j * LS + k], tj);
      if (!(tj > 0.0)) { singular = 1; break; }
      const double ljj = sqrt(tj);
      const double inv = 1.0 / ljj;
      L[j * LS + j] = ljj; S.linv[j] = inv;
      for (int i = j + 1; i < m; ++i) {
        double t = L[i * LS + j];
        for (int k = 0; k < j; ++k) t = fma(-L[i * LS + k], L[j * LS + k], t);
        L[i * LS + j] = t * inv;
      }
    }
    ikRes = 0; id0 = -1; id1 = -1;
    if (singular) {
      ikRes = AFF_CODE_SINGULAR;
    } else {
      // L y = rhs, then L^T z = y (column oriented)
      for (int i = 0; i < m; ++i) {
        const double yfin = rhs[i] * S.linv[i];
        rhs[i] = yfin;
        for (int l = i + 1; l < m; ++l) rhs[l] = fma(-L[l * LS + i], yfin, rhs[l]);
      }
      for (int i = m - 1; i >= 0; --i) {
        const double zfin = rhs[i] * S.linv[i];
        rhs[i] = zfin;
        for (int l = 0; l < i; ++l) rhs[l] = fma(-L[i * LS + l], zfin, rhs[l]);
      }
      // dq, speed limit (:696-723)
      double scale = 1.0;
      for (int j = 0; j < nJ; ++j) {
        double acc = 0.0;
        unsigned mk = S.nz[j];
        while (mk) { const int i = w_ffs(mk) - 1; mk &= mk - 1u; acc = fma(J[i * JS + j], rhs[i], acc); }
        const double dqv = fma(ldg(&P.jwmetric[j]), acc, -S.vbuf[j]);
        S.dqv[j] = dqv;
        const double js = ldg(&P.jspeed[j]);
        if (js < AFFK_NO_LIMIT) { const double a = fabs(dqv), mx = js * dt; if (a > mx) { const double sc = mx / a; if (sc < scale) scale = sc; } }
      }
      const bool scaled = scale < 1.0;
      const double dynLim = -0.99 * qFilt + 1.0;
      const double invdt = 1.0 / dt;
      fast = 0;
      for (int j = 0; j < nJ; ++j) {
        double dqv = S.dqv[j];
        if (scaled) dqv = dqv * (0.99999 * scale);
        double qd = dqv * invdt;
        const double qdPrev = qdot[j];
        const double ja = ldg(&P.jacc[j]);
        if (ja < AFFK_NO_LIMIT) {
          const double lim = (dynLim * ja) * dt;
          const double dv = qd - qdPrev;
          if (dv > lim) qd = qdPrev + lim;
          else if (dv < -lim) qd = qdPrev - lim;
        }
        qdot[j] = qd; q[j] = q[j] + dqv;
        const double js = ldg(&P.jspeed[j]);
        if (js < AFFK_NO_LIMIT && fabs(qd) > js * 1.0) fast = 1;
      }
    }
    }
    T_SYNCI(3);
    T_MARK(2);
    // integration, FK, collision model, checkState (:730-752)
    if (act && !singular) t_fk(c, false);                                                       // ======== phase D
    T_SYNCI(4);
    T_MARK(3);
    if (act && !singular) t_collision(c, r_minDist > 0.001 ? r_minDist : 0.001, &cm);           // ======== phase E
    T_SYNCI(5);
    T_MARK(4);
    if (act && !singular && cm.overflow) r_flags |= AFF_FLAG_LIST_OVERFLOW;
    if (act) {   // ======== phase F: checkState, costs, tracking error
    if (!singular) {
      int viol = 0;
      for (int j = 0; j < nJ; ++j) if (q[j] < ldg(&P.jq_min[j]) || q[j] > ldg(&P.jq_max[j])) ++viol;
      if (fast) ikRes = AFF_CODE_SPEED;
      else if (viol) { ikRes = AFF_CODE_JOINT_LIMIT; id0 = viol; }
      else if (cm.minB1 >= 0 && cm.minDist < 0.001) { ikRes = AFF_CODE_COLLISION; id0 = cm.minB1; id1 = cm.minB2; }
    }
    if (ikRes != 0) {
      if (success) { r_first_code = ikRes; r_first_fail = iter; }
      success = 0; r_code = ikRes; r_fail_step = iter; r_id0 = id0; r_id1 = id1;
    }
    // ---- costs (:226-227); the sum over joints in the order of the 32-lane butterfly of w_tree_sum
    {
      for (int i = 0; i < 16; ++i) {
        double pr2[2];
        for (int h = 0; h < 2; ++h) {
          const int j = i + 16 * h;
          double term = 0.0;
          if (j < nJ) {
            const double jq0 = ldg(&P.jq0[j]);
            const double dqc = q[j] - jq0;
            const double r = (dqc > 0.0) ? ldg(&P.jq_max[j]) - jq0 : jq0 - ldg(&P.jq_min[j]);
            const double u = dqc / r;
            term = (0.5 * ldg(&P.jwjl[j])) * (u * u);
          }
          pr2[h] = term;
        }
        S.tsum[i] = pr2[0] + pr2[1];
      }
      for (int off = 8; off >= 1; off >>= 1) for (int i = 0; i < off; ++i) S.tsum[i] = S.tsum[i] + S.tsum[i + off];
      jlSum = jlSum + S.tsum[0];
      collSum = collSum + cm.cost;
    }
    // ---- geometry of the new state; tracking error (:297-337)
    t_task_geo_all(c, nTasks, tasks);
    if (success) {
      for (int t = 0; t < nTasks; ++t)
        if ((activeMask >> t) & 1u) t_task_dx(c, t, tasks, dxv + S.tdesc[8 * t + TD_XOFF]);
      double ev = -1.0; int ek = 0x7fffffff;
      for (int d = 0; d < xdim; ++d)
        if ((activeMask >> S.taskOfDim[d]) & 1u) { const double a = fabs(dxv[d]); if (a > ev) { ev = a; ek = d; } }
      if (ev > err) {
        err = ev;
        if (err > AFFK_TRACK_MAX) {
          int etask = -1, edim = -1;
          for (int t = 0; t < nTasks; ++t) if (ek >= S.tdesc[8 * t + TD_XOFF]) { etask = t; edim = ek - S.tdesc[8 * t + TD_XOFF]; }
          success = 0; r_code = AFF_CODE_TRACKING; r_first_code = AFF_CODE_TRACKING;
          r_first_fail = iter; r_fail_step = iter; r_id0 = etask; r_id1 = edim;
        }
      }
    }
    if (cm.minDist < r_minDist) { r_minDist = cm.minDist; r_b1 = cm.minB1; r_b2 = cm.minB2; }
    if (qlog && rows < qrows) for (int j = 0; j < nJ; ++j) qlog[(size_t)rows * nJ + j] = q[j];
    rows++;
    if (P.opts.early_exit && !success && successPrev) running = false;     // the reference's commented-out break (:368-373)
    }
    T_MARK(5);
  }
  if (!alive) return;
  aff_result res;
  res.idx = cand; res.success = success; res.code = r_code; res.first_code = r_first_code;
  res.first_fail_step = r_first_fail; res.fail_step = r_fail_step; res.fail_id0 = r_id0; res.fail_id1 = r_id1;
  res.min_dist_b1 = r_b1; res.min_dist_b2 = r_b2; res.n_steps = rows; res.reserved = r_flags;
  res.jmask_bits = (uint64_t)jmaskBits; res.min_dist = r_minDist;
  res.jl_cost = jlSum / (double)rows; res.coll_cost = collSum / (double)rows; res.track_err = err;
  P.results[cand] = res;
}
