#include "plan_build.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <numeric>
#include <thread>
#include <unordered_map>

namespace affk
{

static bool isIdentity12(const double* T)
{
  static const double I[12] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
  for (int i = 0; i < 12; ++i) if (T[i] != I[i]) return false;
  return true;
}

std::string SceneModel::init(const aff_scene_desc* d)
{
  if (!d || d->n_bodies <= 0) return "empty scene description";
  nB = d->n_bodies; nQ = d->n_joints; nS = d->n_shapes;
  parent.assign(d->body_parent, d->body_parent + nB);
  firstJoint.assign(d->body_first_joint, d->body_first_joint + nB);
  nJoints.assign(d->body_n_joints, d->body_n_joints + nB);
  firstShape.assign(d->body_first_shape, d->body_first_shape + nB);
  nShapes.assign(d->body_n_shapes, d->body_n_shapes + nB);
  rbj.assign(d->body_rbj, d->body_rbj + nB);
  A_BP.assign(d->body_A_BP, d->body_A_BP + 12 * (size_t)nB);
  jtype.assign(d->joint_type, d->joint_type + nQ);
  jcon.assign(d->joint_constrained, d->joint_constrained + nQ);
  A_JP.assign(d->joint_A_JP, d->joint_A_JP + 12 * (size_t)nQ);
  qInit.assign(d->joint_q_init, d->joint_q_init + nQ);
  q0.assign(d->joint_q0, d->joint_q0 + nQ);
  qMin.assign(d->joint_q_min, d->joint_q_min + nQ);
  qMax.assign(d->joint_q_max, d->joint_q_max + nQ);
  speed.assign(d->joint_speed_limit, d->joint_speed_limit + nQ);
  acc.assign(d->joint_acc_limit, d->joint_acc_limit + nQ);
  wjl.assign(d->joint_weight_jl, d->joint_weight_jl + nQ);
  wmetric.assign(d->joint_weight_metric, d->joint_weight_metric + nQ);
  stype.assign(d->shape_type, d->shape_type + nS);
  sdist.assign(d->shape_distance, d->shape_distance + nS);
  A_CB.assign(d->shape_A_CB, d->shape_A_CB + 12 * (size_t)nS);
  ext.assign(d->shape_extents, d->shape_extents + 3 * (size_t)nS);
  bpThreshold = d->bp_threshold;
  bpTree.assign(d->bp_tree_root, d->bp_tree_root + d->n_bp_trees);
  bpBody.assign(d->bp_body, d->bp_body + d->n_bp_bodies);
  ns_base = d->ns_base;
  for (int k = 0; k < 2; ++k) { ns_upperarm[k] = d->ns_upperarm[k]; ns_forearm[k] = d->ns_forearm[k]; ns_hand[k] = d->ns_hand[k]; }
  jacobi.assign(nQ, -1); jointBody.assign(nQ, -1);
  nJ = 0;
  for (int j = 0; j < nQ; ++j) if (!jcon[j]) { jacobi[j] = nJ++; ikJoint.push_back(j); }
  for (int b = 0; b < nB; ++b) {
    if (parent[b] >= b) return "scene bodies must list parents before children";
    for (int k = 0; k < nJoints[b]; ++k) jointBody[firstJoint[b] + k] = b;
  }
  explicitBp.assign(nB, 0);
  for (int b : bpBody) { if (b < 0 || b >= nB) return "broad phase body out of range"; explicitBp[b] = 1; }
  for (int b : bpTree) if (b < 0 || b >= nB) return "broad phase tree root out of range";
  if (nJ > AFFK_MAX_NJ) return "more than 32 unconstrained joints";
  if ((int)bpTree.size() > 14) return "more than 14 broad phase trees";
  return "";
}

static int popcount3(int flag) { return (flag & 1) + ((flag >> 1) & 1) + ((flag >> 2) & 1); }

int buildPlan(const SceneModel& S, const aff_task_desc* tasksIn, size_t nTasksIn, const aff_constraint* consIn, size_t nConsIn,
              const aff_candidate* candsIn, size_t nCands, const aff_opts& opts, HostPlan& H, std::string& err)
{
  H = HostPlan();
  KPlan& P = H.plan;
  const int nB = S.nB;
  if (nCands == 0) { err = "empty batch"; return AFF_ERR_INVALID; }

  // ---- objects: bodies that can be re-parented, posed per candidate or toggled
  std::vector<int> objSlotOfBody(nB, -1);
  std::vector<int> objBodies;
  auto addObject = [&](int b) -> bool {
    if (b < 0 || b >= nB) { err = "constraint names a body outside the scene"; return false; }
    if (objSlotOfBody[b] >= 0) return true;
    if (!S.rbj[b] || S.nJoints[b] != 6) { err = "re-parentable / posed / toggled bodies need rigid_body_joints"; return false; }
    if (!isIdentity12(&S.A_BP[12 * b])) { err = "rigid-body-joint objects with an extra body transform are not supported"; return false; }
    if ((int)objBodies.size() >= AFFK_MAX_OBJ) { err = "more than AFFK_MAX_OBJ objects in one batch"; return false; }
    objSlotOfBody[b] = (int)objBodies.size();
    objBodies.push_back(b);
    return true;
  };
  for (size_t i = 0; i < nCands; ++i) {
    const aff_candidate& c = candsIn[i];
    if (c.first_task < 0 || c.n_tasks < 0 || (size_t)(c.first_task + c.n_tasks) > nTasksIn ||
        c.first_constraint < 0 || c.n_constraints < 0 || (size_t)(c.first_constraint + c.n_constraints) > nConsIn) {
      err = "candidate slices outside the task / constraint tables"; return AFF_ERR_INVALID;
    }
    if (c.pose_body >= 0 && !addObject(c.pose_body)) return AFF_ERR_UNSUPPORTED;
    for (int k = 0; k < c.n_constraints; ++k) {
      const aff_constraint& cc = consIn[c.first_constraint + k];
      if ((cc.kind == AFF_CON_CONNECT || cc.kind == AFF_CON_COLLISION) && !addObject(cc.body_a)) return AFF_ERR_UNSUPPORTED;
      if (cc.kind == AFF_CON_CONNECT && (cc.body_b < 0 || cc.body_b >= nB)) { err = "connect parent outside the scene"; return AFF_ERR_INVALID; }
    }
  }
  const int nObj = (int)objBodies.size();

  // ---- dynamic bodies: moved by an IK joint, or an object / below one
  std::vector<char> dynamicB(nB, 0), underObj(nB, 0);
  std::vector<uint32_t> chainB(nB, 0u);
  std::vector<int> treeB(nB, -1);
  for (int b = 0; b < nB; ++b) {
    const int p = S.parent[b];
    uint32_t ch = p >= 0 ? chainB[p] : 0u;
    for (int k = 0; k < S.nJoints[b]; ++k) { const int c = S.jacobi[S.firstJoint[b] + k]; if (c >= 0) ch |= 1u << c; }
    chainB[b] = ch;
    underObj[b] = (objSlotOfBody[b] >= 0) ? (char)(objSlotOfBody[b] + 1) : (p >= 0 ? underObj[p] : 0);
    dynamicB[b] = (ch != 0u) || underObj[b];
    int tr = p >= 0 ? treeB[p] : -1;
    for (size_t t = 0; t < S.bpTree.size(); ++t) if (S.bpTree[t] == b) tr = (int)t;
    treeB[b] = tr;
  }
  // rbj bodies that are not objects: their chains are masked out of the null space for good
  uint32_t rbjStatic = 0u;
  for (int b = 0; b < nB; ++b) if (S.rbj[b] && !underObj[b]) rbjStatic |= chainB[b];

  // ---- which dynamic bodies are needed
  std::vector<char> needed(nB, 0);
  auto need = [&](int b) { for (int a = b; a >= 0 && !needed[a]; a = S.parent[a]) needed[a] = 1; };
  for (int j : S.ikJoint) need(S.jointBody[j]);
  for (int b : objBodies) need(b);
  for (int b : {S.ns_base, S.ns_upperarm[0], S.ns_upperarm[1], S.ns_forearm[0], S.ns_forearm[1], S.ns_hand[0], S.ns_hand[1]}) if (b >= 0) need(b);
  for (size_t i = 0; i < nTasksIn; ++i) {
    const aff_task_desc& t = tasksIn[i];
    for (int b : {t.effector, t.ref_body, t.ref_frame}) {
      if (b >= nB) { err = "task names a body outside the scene"; return AFF_ERR_INVALID; }
      if (b >= 0) need(b);
    }
  }
  for (size_t i = 0; i < nConsIn; ++i) if (consIn[i].kind == AFF_CON_CONNECT && consIn[i].body_b >= 0 && consIn[i].body_b < nB) need(consIn[i].body_b);
  std::vector<char> shapeBody(nB, 0);
  for (int b = 0; b < nB; ++b) {
    bool any = false;
    for (int k = 0; k < S.nShapes[b]; ++k) any = any || S.sdist[S.firstShape[b] + k];
    if (objSlotOfBody[b] >= 0 && S.nShapes[b] > 0) any = true;
    if (any) { shapeBody[b] = 1; need(b); }
  }

  // ---- bodies whose frames the step loop itself reads (task effectors / reference bodies / frames, null-space
  // bodies, objects with everything below them, the bodies objects get connected to) and their ancestors: their
  // nodes come first in the FK schedule, so that the block-per-candidate kernel can leave the rest (fingers, the
  // parts of the robot no task refers to: only the collision model reads those) to a helper warp.
  std::vector<char> critB(nB, 0);
  {
    auto crit = [&](int b) { for (int a = b; a >= 0 && !critB[a]; a = S.parent[a]) critB[a] = 1; };
    for (int b : {S.ns_base, S.ns_upperarm[0], S.ns_upperarm[1], S.ns_forearm[0], S.ns_forearm[1], S.ns_hand[0], S.ns_hand[1]}) if (b >= 0) crit(b);
    for (size_t i = 0; i < nTasksIn; ++i) for (int b : {tasksIn[i].effector, tasksIn[i].ref_body, tasksIn[i].ref_frame}) if (b >= 0 && b < nB) crit(b);
    for (size_t i = 0; i < nConsIn; ++i) if (consIn[i].kind == AFF_CON_CONNECT && consIn[i].body_b >= 0 && consIn[i].body_b < nB) crit(consIn[i].body_b);
    for (int b : objBodies) crit(b);
    for (int b = 0; b < nB; ++b) if (underObj[b]) crit(b);
  }
  std::vector<char> nodeCrit;     // per dynamic node

  // ---- nodes
  std::vector<int> bodyRef(nB, AFFK_WORLD), jointRef(S.nQ, AFFK_WORLD);
  H.j_node.assign(S.nJ, -1); H.j_type.assign(S.nJ, 0);
  auto mkNode = [](int parentRef, int kind, const double* T) {
    KNode n; std::memset(&n, 0, sizeof(n));
    n.parent = parentRef; n.kind = kind; n.tree = -1;
    std::memcpy(n.T, T, sizeof(double) * 12);
    n.identityT = isIdentity12(T) ? 1 : 0;
    return n;
  };
  for (int b = 0; b < nB; ++b) {
    const int pRef = S.parent[b] >= 0 ? bodyRef[S.parent[b]] : AFFK_WORLD;
    if (!dynamicB[b]) {
      int cur = pRef;
      for (int k = 0; k < S.nJoints[b]; ++k) {
        const int j = S.firstJoint[b] + k;
        KNode n = mkNode(cur, AFFK_NODE_JOINT_CONST, &S.A_JP[12 * j]);
        n.jtype = S.jtype[j]; n.qconst = S.qInit[j];
        H.stat.push_back(n);
        cur = AFFK_STATIC_REF((int)H.stat.size() - 1);
        jointRef[j] = cur;
      }
      if (!isIdentity12(&S.A_BP[12 * b]) || cur == AFFK_WORLD) {
        H.stat.push_back(mkNode(cur, AFFK_NODE_FIXED, &S.A_BP[12 * b]));
        cur = AFFK_STATIC_REF((int)H.stat.size() - 1);
      }
      bodyRef[b] = cur;
      continue;
    }
    if (!needed[b]) continue;
    const int uo = underObj[b];
    if (objSlotOfBody[b] >= 0) {
      static const double I[12] = {0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0, 1};
      KNode n = mkNode(pRef, AFFK_NODE_OBJECT, I);
      n.qidx = objSlotOfBody[b]; n.under_obj = uo; n.chain = 0u; n.tree = -1;
      H.dyn.push_back(n); nodeCrit.push_back(1);
      bodyRef[b] = (int)H.dyn.size() - 1;
      continue;
    }
    int cur = pRef;
    for (int k = 0; k < S.nJoints[b]; ++k) {
      const int j = S.firstJoint[b] + k;
      const int c = S.jacobi[j];
      KNode n = mkNode(cur, c >= 0 ? AFFK_NODE_JOINT_IK : AFFK_NODE_JOINT_CONST, &S.A_JP[12 * j]);
      n.jtype = S.jtype[j]; n.qconst = S.qInit[j]; n.qidx = c >= 0 ? c : 0; n.under_obj = uo;
      // chain up to and including this joint
      uint32_t ch = S.parent[b] >= 0 ? chainB[S.parent[b]] : 0u;
      for (int kk = 0; kk <= k; ++kk) { const int cc = S.jacobi[S.firstJoint[b] + kk]; if (cc >= 0) ch |= 1u << cc; }
      n.chain = uo ? 0u : ch;
      n.tree = uo ? -1 : treeB[b];
      H.dyn.push_back(n); nodeCrit.push_back(critB[b]);
      cur = (int)H.dyn.size() - 1;
      jointRef[j] = cur;
      if (c >= 0) { H.j_node[c] = cur; H.j_type[c] = S.jtype[j]; }
    }
    bool isTreeRoot = false;
    for (int r : S.bpTree) isTreeRoot = isTreeRoot || (r == b);
    // The body frame aliases its last joint frame (or its parent) when A_BP is the identity.
    const bool alias = isIdentity12(&S.A_BP[12 * b]) && cur >= 0 && !(isTreeRoot && S.nJoints[b] == 0);
    if (!alias) {
      KNode n = mkNode(cur, AFFK_NODE_FIXED, &S.A_BP[12 * b]);
      n.under_obj = uo; n.chain = uo ? 0u : chainB[b]; n.tree = uo ? -1 : treeB[b];
      H.dyn.push_back(n); nodeCrit.push_back(critB[b]);
      cur = (int)H.dyn.size() - 1;
    }
    bodyRef[b] = cur;
  }
  if ((int)H.dyn.size() > AFFK_MAX_DYN) { err = "too many dynamic nodes (" + std::to_string(H.dyn.size()) + ")"; return AFF_ERR_CAPACITY; }
  for (int c = 0; c < S.nJ; ++c) if (H.j_node[c] < 0) { err = "internal: IK joint without a node"; return AFF_ERR_INVALID; }

  // ---- FK schedule: two nodes per round, parents in earlier rounds
  {
    const int nD = (int)H.dyn.size();
    std::vector<int> roundOf(nD, 0), fill;
    // two passes: the nodes the step loop reads (a node's parent is one of them if the node is), then the rest,
    // which may fill free places of the early rounds
    int nCritRounds = 0;
    for (int pass = 0; pass < 2; ++pass) {
      for (int n = 0; n < nD; ++n) {
        if ((nodeCrit[n] != 0) != (pass == 0)) continue;
        int r = 0;
        const int p = H.dyn[n].parent;
        if (p >= 0) r = roundOf[p] + 1;
        // an object may later hang on any of the nodes objects get connected to (all in the first pass; a
        // parent the schedule still computes later makes KPlan::warp_ok zero): after everything earlier that is not below an object
        if (H.dyn[n].kind == AFFK_NODE_OBJECT) for (int k = 0; k < n; ++k) if (!H.dyn[k].under_obj && nodeCrit[k]) r = std::max(r, roundOf[k] + 1);
        while (true) {
          if ((int)fill.size() <= r) fill.resize(r + 1, 0);
          if (fill[r] < 2) break;
          ++r;
        }
        roundOf[n] = r; fill[r]++;
      }
      if (pass == 0) nCritRounds = (int)fill.size();
    }
    P.nRoundsCrit = nCritRounds;
    const int nR = (int)fill.size();
    if (nR > AFFK_MAX_ROUNDS) { err = "FK schedule too deep"; return AFF_ERR_CAPACITY; }
    H.rounds.assign(2 * (size_t)nR, -1);
    std::vector<int> used(nR, 0);
    for (int n = 0; n < nD; ++n) { const int r = roundOf[n]; H.rounds[2 * r + used[r]++] = n; }
    P.nRounds = nR;
    H.roundOf.assign(roundOf.begin(), roundOf.end());
  }
  // ---- static parents get a shared-memory slot behind the dynamic nodes; FK program words
  std::vector<int> statParSlot;   // static node index -> slot, built lazily
  auto slotOfRef = [&](int ref) -> int {
    if (ref >= 0) return ref;
    for (size_t k = 0; k < H.statpar_ref.size(); ++k) if (H.statpar_ref[k] == ref) return (int)H.dyn.size() + (int)k;
    H.statpar_ref.push_back(ref);
    return (int)H.dyn.size() + (int)H.statpar_ref.size() - 1;
  };
  {
    const int nD = (int)H.dyn.size();
    H.dynT.resize(12 * (size_t)std::max(nD, 1)); H.dynQ.resize((size_t)std::max(nD, 1));
    for (int n = 0; n < nD; ++n) { std::memcpy(&H.dynT[12 * n], H.dyn[n].T, sizeof(double) * 12); H.dynQ[n] = H.dyn[n].qconst; }
    // per-lane FK table
    const int nR = (int)H.rounds.size() / 2;
    KFkEntry idle; std::memset(&idle, 0, sizeof(idle));
    H.fktab.assign((size_t)(nR + 1) * 32, idle);   // idle lanes get their destination (the spare element) in patchFkAux; one spare round: the prefetch of k_fk_t<1> reads one round ahead
    P.fk_obj_rounds[0] = P.fk_obj_rounds[1] = P.fk_const_rounds[0] = P.fk_const_rounds[1] = 0u;
    for (int r = 0; r < nR; ++r)
      for (int grp = 0; grp < 2; ++grp) {
        const int n = H.rounds[2 * r + grp];
        if (n < 0) continue;
        const KNode& nd = H.dyn[n];
        const int pslot = slotOfRef(nd.parent);
        for (int e = 0; e < 12; ++e) {
          const int lane = grp * 12 + e;
          KFkEntry& en = H.fktab[(size_t)r * 32 + lane];
          const int row = e < 3 ? -1 : (e - 3) / 3, col = e < 3 ? e : (e - 3) % 3;
          const int tb = e < 3 ? 0 : 3 + 3 * row;
          en.t0 = nd.T[tb]; en.t1 = nd.T[tb + 1]; en.t2 = nd.T[tb + 2];
          // aux (w1 bits 16-27): index of the lane's extra operand, filled in once the layout is known
          // (patchFkAux below); until then it holds the element relative to the node array for COPY
          uint32_t op = 2, partner = (uint32_t)lane, pbase = (uint32_t)(12 * pslot), dst = (uint32_t)(12 * n + e), w1 = 0, aux = 0;
          if (nd.kind == AFFK_NODE_OBJECT) { op = 6; pbase = 0; w1 = (uint32_t)(nd.qidx & 3) | ((uint32_t)n << 8); P.fk_obj_rounds[(r >> 5) & 1] |= 1u << (r & 31); }
          else if (nd.kind == AFFK_NODE_FIXED) { if (nd.identityT) { op = 1; pbase = 0; aux = (uint32_t)(12 * pslot + e); } }
          else {
            const bool isRot = nd.jtype >= AFF_JNT_ROT_X;
            const int d = isRot ? nd.jtype - AFF_JNT_ROT_X : nd.jtype;
            if (isRot && e >= 3) {
              const int a = (d + 1) % 3, b = (d + 2) % 3;
              if (row == a) { op = 3; partner = (uint32_t)(grp * 12 + 3 + 3 * b + col); }
              else if (row == b) { op = 4; partner = (uint32_t)(grp * 12 + 3 + 3 * a + col); }
            } else if (!isRot && e < 3) { op = 5; partner = (uint32_t)(grp * 12 + 3 + 3 * d + e); }
            // only the lanes that apply the joint carry its index; the others stay plain compositions
            if (op != 2) {
              w1 = (nd.kind == AFFK_NODE_JOINT_IK) ? (uint32_t)(nd.qidx & 63) : (0x80u | ((uint32_t)n << 8));
              if (nd.kind != AFFK_NODE_JOINT_IK) P.fk_const_rounds[(r >> 5) & 1] |= 1u << (r & 31);
            }
          }
          if (aux > 4095u) { err = "FK table: node array too large"; return AFF_ERR_CAPACITY; }
          w1 |= aux << 16;
          if (pbase > 4095u || dst > 4095u) { err = "FK table: node array too large"; return AFF_ERR_CAPACITY; }
          en.w0 = pbase | (dst << 12) | (op << 24) | (partner << 27);
          en.w1 = w1;
        }
      }
  }
  // parent slot of every dynamic node (thread-per-candidate FK walks the nodes in index order)
  H.dyn_pslot.resize(H.dyn.size());
  for (size_t n = 0; n < H.dyn.size(); ++n) H.dyn_pslot[n] = slotOfRef(H.dyn[n].parent);
  // objects below objects (a connect parent under another object) would need a dynamic schedule
  for (int b : objBodies) if (S.parent[b] >= 0 && underObj[S.parent[b]]) { err = "objects nested in objects are not supported"; return AFF_ERR_UNSUPPORTED; }

  // ---- shapes
  auto addShapes = [&](int b, bool dynamicSide) {
    KShapeBody sb; std::memset(&sb, 0, sizeof(sb));
    sb.body = b; sb.node = bodyRef[b]; sb.explicit_bp = S.explicitBp[b];
    sb.obj_slot = underObj[b]; sb.toggle_slot = objSlotOfBody[b] >= 0 ? objSlotOfBody[b] + 1 : 0;
    sb.tree = underObj[b] ? -1 : treeB[b];
    std::vector<KShape>& dst = dynamicSide ? H.dshape : H.sshape;
    sb.first_shape = (int)dst.size();
    for (int k = 0; k < S.nShapes[b]; ++k) {
      const int s = S.firstShape[b] + k;
      if (!S.sdist[s] && !sb.toggle_slot) continue;   // can never become active
      KShape sh; std::memset(&sh, 0, sizeof(sh));
      sh.type = S.stype[s]; sh.active0 = S.sdist[s]; sh.node = bodyRef[b];
      std::memcpy(sh.ext, &S.ext[3 * s], sizeof(double) * 3);
      std::memcpy(sh.A_CB, &S.A_CB[12 * s], sizeof(double) * 12);
      // bounding radius about the shape centre (segment midpoint / frame origin); used for culling with a margin only
      if (sh.type == AFF_SHAPE_SSL) sh.bound = 0.5 * std::fabs(sh.ext[2]) + sh.ext[0];
      else if (sh.type == AFF_SHAPE_SPHERE) sh.bound = sh.ext[0];
      else if (sh.type == AFF_SHAPE_BOX) sh.bound = 0.5 * std::sqrt(sh.ext[0] * sh.ext[0] + sh.ext[1] * sh.ext[1] + sh.ext[2] * sh.ext[2]);
      else sh.bound = std::sqrt(sh.ext[0] * sh.ext[0] + 0.25 * sh.ext[2] * sh.ext[2]);
      sh.bound *= 1.0 + 1.0e-12;
      if (sh.active0) sb.any_active0 = 1;
      dst.push_back(sh);
    }
    sb.n_shapes = (int)dst.size() - sb.first_shape;
    if (sb.n_shapes > 0) (dynamicSide ? H.dbody : H.sbody).push_back(sb);
  };
  // dynamic bodies in body order; static bodies with the ones the per-step AABB test looks at first
  // (not explicit, some shape active from the start), so that the test loops over a prefix only
  auto nearCandidate = [&](int b) {
    if (S.explicitBp[b]) return false;
    for (int k = 0; k < S.nShapes[b]; ++k) if (S.sdist[S.firstShape[b] + k]) return true;
    return false;
  };
  for (int b = 0; b < nB; ++b) if (shapeBody[b] && dynamicB[b]) addShapes(b, true);
  for (int b = 0; b < nB; ++b) if (shapeBody[b] && !dynamicB[b] && nearCandidate(b)) addShapes(b, false);
  P.nStatNear = (int)H.sbody.size();
  for (int b = 0; b < nB; ++b) if (shapeBody[b] && !dynamicB[b] && !nearCandidate(b)) addShapes(b, false);
  if ((int)H.dshape.size() > AFFK_MAX_DSHAPE || (int)H.dbody.size() > AFFK_MAX_DBODY) { err = "too many dynamic shapes"; return AFF_ERR_CAPACITY; }
  // dynamic shapes by the node they hang on (CSR): the thread-per-candidate FK writes their world records
  {
    const int nD = (int)H.dyn.size();
    H.node_shape_off.assign((size_t)nD + 1, 0);
    for (const KShape& sh : H.dshape) if (sh.node >= 0 && sh.node < nD) H.node_shape_off[(size_t)sh.node + 1]++;
    for (int n = 0; n < nD; ++n) H.node_shape_off[(size_t)n + 1] += H.node_shape_off[(size_t)n];
    H.node_shape_idx.assign(H.dshape.size() + 1, 0);
    std::vector<int32_t> fill(H.node_shape_off.begin(), H.node_shape_off.end() - 1);
    for (size_t si = 0; si < H.dshape.size(); ++si) { const int n = H.dshape[si].node; if (n >= 0 && n < nD) H.node_shape_idx[(size_t)fill[(size_t)n]++] = (int32_t)si; }
    for (const KShape& sh : H.dshape) if (sh.node < 0 || sh.node >= nD) { err = "internal: dynamic shape on a static node"; return AFF_ERR_INVALID; }
  }
  // ---- precompiled shape pairs of the collision model, by type.  A body pair is listed when it can
  // ever be live without the AABB test: two dynamic bodies that are (or may come to be) in different
  // trees or one of which is explicit / an object, and dynamic tree bodies against explicit static bodies.
  {
    auto isSeg = [](int t) { return t == AFF_SHAPE_SSL || t == AFF_SHAPE_SPHERE; };
    auto metaOf = [&](const KShapeBody& A, int idxA, bool actA, const KShapeBody& B, int idxB, bool actB, bool swap) {
      uint32_t m = 0;
      m |= (uint32_t)(A.toggle_slot ? A.toggle_slot : A.obj_slot) & 3u;
      m |= ((uint32_t)(B.toggle_slot ? B.toggle_slot : B.obj_slot) & 3u) << 2;
      m |= (swap ? 1u : 0u) << 4;
      m |= ((uint32_t)(A.tree + 1) & 15u) << 9;
      m |= ((uint32_t)(B.tree + 1) & 15u) << 13;
      m |= (A.explicit_bp ? 1u : 0u) << 17;
      m |= (B.explicit_bp ? 1u : 0u) << 18;
      m |= (actA ? 1u : 0u) << 19;
      m |= (actB ? 1u : 0u) << 20;
      m |= ((uint32_t)idxA & 31u) << 21;
      m |= ((uint32_t)idxB & 31u) << 26;
      // fixed membership and always paired: both in (different) trees, or one in a tree and the other explicit
      const bool objInvolved = A.obj_slot || B.obj_slot;
      const bool always = !objInvolved && ((A.tree >= 0 && B.tree >= 0) || (A.tree >= 0 && B.explicit_bp) || (B.tree >= 0 && A.explicit_bp));
      if (!always) m |= 1u << 31;
      return (int32_t)m;
    };
    // body-level candidates: (dynamic body index, other: >= 0 dynamic, < 0 static)
    auto addBodyPair = [&](int ia, int ib) {
      const KShapeBody& A = H.dbody[ia];
      const KShapeBody& B = ib >= 0 ? H.dbody[ib] : H.sbody[-1 - ib];
      const std::vector<KShape>& shA = H.dshape;
      const std::vector<KShape>& shB = ib >= 0 ? H.dshape : H.sshape;
      const bool aFirst = A.body < B.body;      // canonical argument order: ascending scene body id
      const int key = aFirst ? ((A.body << 15) | B.body) : ((B.body << 15) | A.body);
      for (int k1 = 0; k1 < A.n_shapes; ++k1)
        for (int k2 = 0; k2 < B.n_shapes; ++k2) {
          const KShape& s1 = shA[A.first_shape + k1];
          const KShape& s2 = shB[B.first_shape + k2];
          const int i1 = A.first_shape + k1;
          const int i2 = ib >= 0 ? B.first_shape + k2 : -1 - (B.first_shape + k2);
          KPair pr;
          pr.key = key;
          if (isSeg(s1.type) && isSeg(s2.type)) {
            pr.a = i1; pr.b = i2;
            pr.meta = metaOf(A, ia, s1.active0 != 0, B, ib >= 0 ? ib : 31, s2.active0 != 0, !aFirst);
            H.pairsSS.push_back(pr);
          } else if (isSeg(s1.type)) {
            pr.a = i1; pr.b = i2;
            pr.meta = metaOf(A, ia, s1.active0 != 0, B, ib >= 0 ? ib : 31, s2.active0 != 0, false);
            H.pairsSB.push_back(pr);
          } else if (isSeg(s2.type)) {
            pr.a = i2; pr.b = i1;      // the segment goes first; meta keeps A/B = body roles of (a, b)
            pr.meta = metaOf(B, ib >= 0 ? ib : 31, s2.active0 != 0, A, ia, s1.active0 != 0, false);
            H.pairsSB.push_back(pr);
          }
          // box/cylinder against box/cylinder: no distance function, never part of any result
        }
    };
    for (size_t i = 0; i < H.dbody.size(); ++i) {
      const KShapeBody& a = H.dbody[i];
      const bool aTree = a.tree >= 0 || a.obj_slot;
      for (size_t j = i + 1; j < H.dbody.size(); ++j) {
        const KShapeBody& b = H.dbody[j];
        const bool bTree = b.tree >= 0 || b.obj_slot;
        if (!aTree && !bTree) continue;
        if (a.tree >= 0 && b.tree >= 0 && a.tree == b.tree && !a.obj_slot && !b.obj_slot) continue;
        addBodyPair((int)i, (int)j);
      }
      if (aTree) for (size_t j = 0; j < H.sbody.size(); ++j) if (H.sbody[j].explicit_bp && H.sbody[j].any_active0) addBodyPair((int)i, -1 - (int)j);
    }
    if (H.pairsSS.size() > AFFK_MAX_PAIRS || H.pairsSB.size() > AFFK_MAX_PAIRS) { err = "too many precompiled collision pairs"; return AFF_ERR_CAPACITY; }
    // Pairs of the same storage kind (partner in shared memory / in the static tables) and the same need
    // for a per-step liveness decision sit next to each other, so that most rounds of 32 take one side of
    // those branches as a whole warp.  The order of the lists has no effect on the results (keyed minimum,
    // costs summed in key order).
    auto kindOf = [](const KPair& p) { return ((p.a >= 0) ? 0 : 1) | ((p.b >= 0) ? 0 : 2) | (((uint32_t)p.meta >> 31) ? 4 : 0); };
    // within a kind, pairs of the same dynamic shape `a` follow each other (the thread-per-candidate kernel keeps its record in registers)
    std::stable_sort(H.pairsSS.begin(), H.pairsSS.end(), [&](const KPair& x, const KPair& y) { return kindOf(x) != kindOf(y) ? kindOf(x) < kindOf(y) : x.a < y.a; });
    // segment-box before segment-cylinder (the cylinder distance is an iteration: keep it out of box rounds)
    auto kindSB = [&](const KPair& p) { const int t = (p.b >= 0) ? H.dshape[p.b].type : H.sshape[-1 - p.b].type; return kindOf(p) | (t == AFF_SHAPE_CYLINDER ? 8 : 0); };
    std::stable_sort(H.pairsSB.begin(), H.pairsSB.end(), [&](const KPair& x, const KPair& y) { return kindSB(x) != kindSB(y) ? kindSB(x) < kindSB(y) : x.a < y.a; });
    if (H.dbody.size() > 31) { err = "too many dynamic shape bodies"; return AFF_ERR_CAPACITY; }
  }

  // ---- joints (jacobi order)
  for (int c = 0; c < S.nJ; ++c) {
    const int j = S.ikJoint[c];
    H.jq_init.push_back(S.qInit[j]); H.jq0.push_back(S.q0[j]); H.jq_min.push_back(S.qMin[j]); H.jq_max.push_back(S.qMax[j]);
    H.jspeed.push_back(S.speed[j]); H.jacc.push_back(S.acc[j]); H.jwjl.push_back(S.wjl[j]); H.jwmetric.push_back(S.wmetric[j]);
  }
  for (int s = 0; s < nObj; ++s) {
    const int b = objBodies[s];
    H.obj_node.push_back(bodyRef[b]); H.obj_body.push_back(b);
    H.obj_parent_slot.push_back(slotOfRef(H.dyn[bodyRef[b]].parent));
    H.obj_parent_tree.push_back(S.parent[b] >= 0 ? treeB[S.parent[b]] : -1);
    for (int k = 0; k < 6; ++k) H.obj_q6.push_back(S.qInit[S.firstJoint[b] + k]);
  }

  // ---- tasks.  Candidates with byte-identical task slices (every perturbed candidate of one command and
  // hand) share one copy: less to upload, and the lanes of a warp of the thread-per-candidate kernel
  // read the same addresses (one broadcast load instead of 32).
  std::vector<int32_t> taskBaseOf(nCands, 0);
  std::vector<aff_task_desc> tasksUniq;
  {
    std::unordered_map<std::string, int32_t> seen;
    int32_t lastIn = -1, lastN = -1, lastOut = -1;
    for (size_t i = 0; i < nCands; ++i) {
      const aff_candidate& c = candsIn[i];
      if (lastOut >= 0 && c.n_tasks == lastN &&
          (c.first_task == lastIn || std::memcmp(tasksIn + c.first_task, tasksIn + lastIn, sizeof(aff_task_desc) * (size_t)c.n_tasks) == 0)) {
        taskBaseOf[i] = lastOut;
        continue;
      }
      std::string key((const char*)(tasksIn + c.first_task), sizeof(aff_task_desc) * (size_t)c.n_tasks);
      auto it = seen.find(key);
      if (it == seen.end()) {
        it = seen.emplace(std::move(key), (int32_t)tasksUniq.size()).first;
        tasksUniq.insert(tasksUniq.end(), tasksIn + c.first_task, tasksIn + c.first_task + c.n_tasks);
      }
      taskBaseOf[i] = it->second;
      lastIn = c.first_task; lastN = c.n_tasks; lastOut = it->second;
    }
  }
  const size_t nTasksU = tasksUniq.size();
  H.tasks.resize(nTasksU);
  for (size_t i = 0; i < nTasksU; ++i) {
    const aff_task_desc& t = tasksUniq[i];
    KTask& k = H.tasks[i];
    std::memset(&k, 0, sizeof(k));
    if (t.kind < AFF_TASK_XYZ || t.kind > AFF_TASK_JOINTS) { err = "unsupported task kind"; return AFF_ERR_UNSUPPORTED; }
    k.kind = t.kind; k.axis = (t.axis >= 0 && t.axis < 3) ? t.axis : 2;
    k.eff = t.effector >= 0 ? slotOfRef(bodyRef[t.effector]) : -1;
    k.ref = t.ref_body >= 0 ? slotOfRef(bodyRef[t.ref_body]) : -1;
    k.frame = (t.ref_frame == -2) ? -1 : (t.ref_frame >= 0 ? slotOfRef(bodyRef[t.ref_frame]) : k.ref);
    if (t.kind != AFF_TASK_JOINTS && t.effector < 0) { err = "task without effector"; return AFF_ERR_INVALID; }
    // a reference to the world origin itself cannot be told apart from "no reference": forbid bodies aliased to world
    if (t.kind == AFF_TASK_JOINTS) {
      if (t.n_joints < 0 || t.n_joints > AFF_MAX_TASK_JOINTS) { err = "bad joint count in Joints task"; return AFF_ERR_INVALID; }
      k.n_joints = t.n_joints;
      for (int q = 0; q < t.n_joints; ++q) {
        const int j = t.joints[q];
        if (j < 0 || j >= S.nQ) { err = "Joints task names a joint outside the scene"; return AFF_ERR_INVALID; }
        k.jidx[q] = S.jacobi[j] >= 0 ? S.jacobi[j] : -1;
        k.jconst[q] = S.qInit[j];
      }
    }
    k.has_region = t.has_region; k.region_min = t.region_min; k.region_max = t.region_max; k.region_dx_scaling = t.region_dx_scaling;
  }

  // ---- candidates
  P.warp_ok = 1;
  for (int b : objBodies) { const int pr = H.dyn[bodyRef[b]].parent; if (pr >= 0 && H.roundOf[pr] >= H.roundOf[bodyRef[b]]) P.warp_ok = 0; }
  H.cands.resize(nCands);
  H.vias.reserve(nConsIn); H.acts.reserve(nConsIn);
  int maxViaUnk = 0;
  // task row offsets once per shared task slice (the candidates only read them)
  {
    std::vector<char> done(H.tasks.size() + 1, 0);
    for (size_t i = 0; i < nCands; ++i) {
      const int base = taskBaseOf[i];
      if (candsIn[i].n_tasks > AFFK_MAX_TASKS) { err = "too many tasks in one candidate"; return AFF_ERR_CAPACITY; }
      if (candsIn[i].n_tasks == 0 || done[(size_t)base]) continue;
      done[(size_t)base] = 1;
      int xdim = 0;
      for (int t = 0; t < candsIn[i].n_tasks; ++t) {
        KTask& k = H.tasks[(size_t)base + t];
        k.xoff = xdim;
        xdim += (k.kind == AFF_TASK_XYZ) ? 3 : (k.kind == AFF_TASK_POLAR ? 2 : (k.kind == AFF_TASK_JOINTS ? k.n_joints : 1));
      }
    }
  }
  // static frames a connect can name get their slot now (slotOfRef appends on first sight: not for the workers)
  for (size_t i = 0; i < nCands; ++i)
    for (int k = 0; k < candsIn[i].n_constraints; ++k) {
      const aff_constraint& cc = consIn[candsIn[i].first_constraint + k];
      if (cc.kind == AFF_CON_CONNECT && cc.body_b >= 0 && cc.body_b < nB) (void)slotOfRef(bodyRef[cc.body_b]);
    }
  // The per-candidate tables are built by a few host threads over contiguous candidate ranges (each into its own
  // vectors, merged in order afterwards: the result does not depend on the number of threads).
  struct Local
  {
    std::vector<KVia> vias; std::vector<KAct> acts; std::vector<KGraphCon> gcons;
    std::vector<std::pair<int, int>> order;
    int maxXdim = 0, maxSteps = 0, maxViaUnk = 0, warpOk = 1, rc = AFF_OK;
    std::string err;
  };
  auto work = [&](size_t lo, size_t hi, Local& L) -> int {
    for (size_t i = lo; i < hi; ++i) {
      const aff_candidate& c = candsIn[i];
      KCand& kc = H.cands[i];
      std::memset(&kc, 0, sizeof(kc));
      if (c.n_tasks > AFFK_MAX_TASKS) { L.err = "too many tasks in one candidate"; return AFF_ERR_CAPACITY; }
      kc.first_task = taskBaseOf[i]; kc.n_tasks = c.n_tasks;
      int xdim = 0;
      int toff[AFFK_MAX_TASKS + 1];
      for (int t = 0; t < c.n_tasks; ++t) {
        const KTask& k = H.tasks[kc.first_task + t];
        toff[t] = xdim;      // = k.xoff, set once per shared slice before the workers start
        xdim += (k.kind == AFF_TASK_XYZ) ? 3 : (k.kind == AFF_TASK_POLAR ? 2 : (k.kind == AFF_TASK_JOINTS ? k.n_joints : 1));
      }
      toff[c.n_tasks] = xdim;
      if (xdim > AFFK_MAX_XDIM) { L.err = "candidate has more than AFFK_MAX_XDIM task dimensions"; return AFF_ERR_CAPACITY; }
      kc.xdim = xdim;
      L.maxXdim = std::max(L.maxXdim, xdim);
      const aff_constraint* cc = consIn + c.first_constraint;
      double endT = 0.0, startT = 1.0e308;
      for (int k = 0; k < c.n_constraints; ++k) { if (cc[k].t > endT) endT = cc[k].t; if (cc[k].t < startT) startT = cc[k].t; }
      kc.end_abs = endT;
      kc.start_abs = (c.n_constraints == 0) ? 0.0 : (startT < 0.0 ? 0.0 : startT);
      kc.n_steps = (int)std::lround(endT / opts.dt) + opts.n_after_steps;
      L.maxSteps = std::max(L.maxSteps, kc.n_steps);
      // vias by (dimension, time), stable in insertion order
      kc.first_via = (int)L.vias.size();
      int unknowns[AFFK_MAX_XDIM];   // per dimension, whole trajectory
      for (int d = 0; d < xdim; ++d) unknowns[d] = 0;
      L.order.clear();
      for (int k = 0; k < c.n_constraints; ++k) {
        if (cc[k].kind != AFF_CON_VIA) continue;
        if (cc[k].task < 0 || cc[k].task >= c.n_tasks || cc[k].dim < 0 || toff[cc[k].task] + cc[k].dim >= toff[cc[k].task + 1]) {
          L.err = "via point outside its task"; return AFF_ERR_INVALID;
        }
        L.order.emplace_back(toff[cc[k].task] + cc[k].dim, k);
      }
      std::stable_sort(L.order.begin(), L.order.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
        if (a.first != b.first) return a.first < b.first;
        return cc[a.second].t < cc[b.second].t;
      });
      {
        size_t oi = 0;
        for (int d = 0; d <= xdim; ++d) {
          kc.via_off[d] = (int)(L.vias.size() - kc.first_via);
          while (d < xdim && oi < L.order.size() && L.order[oi].first == d) {
            const aff_constraint& v = cc[L.order[oi].second];
            KVia kv; kv.t = v.t; kv.x = v.x; kv.xd = v.xd; kv.xdd = v.xdd; kv.flag = v.flag & 7; kv.pad = 0;
            unknowns[d] += popcount3(kv.flag);
            L.vias.push_back(kv);
            ++oi;
          }
        }
      }
      // Largest window any step can see: with constraint i the next one to come (tnow <= t_i), the
      // window ends at the first full constraint at least one horizon later (k_via_step).
      for (int d = 0; d < xdim; ++d) {
        const KVia* v = L.vias.data() + kc.first_via + kc.via_off[d];
        const int n = kc.via_off[d + 1] - kc.via_off[d];
        for (int i = 0; i < n; ++i) {
          int G = -1, lastFull = -1;
          for (int j = i; j < n; ++j) if (v[j].flag == 7) { lastFull = j; if (v[j].t - v[i].t >= AFFK_HORIZON) { G = j; break; } }
          if (G < 0) G = lastFull >= 0 ? lastFull : n - 1;
          int K = 0;
          for (int j = i; j <= G; ++j) K += popcount3(v[j].flag);
          L.maxViaUnk = std::max(L.maxViaUnk, K);
        }
      }
      if (L.maxViaUnk > AFFK_MAX_VIA_UNK) { L.err = "a trajectory window has more than 12 constrained values"; return AFF_ERR_CAPACITY; }
      // activations by (task, time)
      kc.first_act = (int)L.acts.size();
      L.order.clear();
      for (int k = 0; k < c.n_constraints; ++k) {
        if (cc[k].kind != AFF_CON_ACTIVATION) continue;
        if (cc[k].task < 0 || cc[k].task >= c.n_tasks) { L.err = "activation outside the task list"; return AFF_ERR_INVALID; }
        L.order.emplace_back(cc[k].task, k);
      }
      std::stable_sort(L.order.begin(), L.order.end(), [&](const std::pair<int, int>& a, const std::pair<int, int>& b) {
        if (a.first != b.first) return a.first < b.first;
        return cc[a.second].t < cc[b.second].t;
      });
      {
        size_t oi = 0;
        for (int t = 0; t <= c.n_tasks; ++t) {
          kc.act_off[t] = (int)(L.acts.size() - kc.first_act);
          while (t < c.n_tasks && oi < L.order.size() && L.order[oi].first == t) {
            const aff_constraint& a = cc[L.order[oi].second];
            KAct ka; ka.t = a.t; ka.horizon = a.horizon; ka.on = a.flag ? 1 : 0; ka.pad = 0;
            L.acts.push_back(ka);
            ++oi;
          }
        }
      }
      // graph constraints in list order
      kc.first_gcon = (int)L.gcons.size();
      for (int k = 0; k < c.n_constraints; ++k) {
        if (cc[k].kind != AFF_CON_CONNECT && cc[k].kind != AFF_CON_COLLISION) continue;
        KGraphCon g; g.t = cc[k].t; g.kind = cc[k].kind; g.on = (cc[k].kind == AFF_CON_CONNECT) ? (cc[k].dim == 1 ? 1 : 0) : (cc[k].flag ? 1 : 0); g.slot = objSlotOfBody[cc[k].body_a];
        g.parent = (cc[k].kind == AFF_CON_CONNECT) ? bodyRef[cc[k].body_b] : AFFK_WORLD;
        g.parent_slot = (cc[k].kind == AFF_CON_CONNECT) ? slotOfRef(g.parent) : 0;
        g.parent_tree = -1;
        if (cc[k].kind == AFF_CON_CONNECT) g.parent_tree = underObj[cc[k].body_b] ? -2 - (underObj[cc[k].body_b] - 1) : treeB[cc[k].body_b];
        if (cc[k].kind == AFF_CON_CONNECT && cc[k].body_b >= 0 && underObj[cc[k].body_b] == g.slot + 1) { L.err = "connect onto the object's own subtree"; return AFF_ERR_INVALID; }
        // The warp-per-candidate FK composes an object with its parent's pose of the same step only if the
        // parent sits in an earlier round of the schedule; a parent in a later round (an object listed before
        // the robot, or put onto another object's subtree) would be read one step late.  Such batches run on
        // the thread-per-candidate kernel, whose FK walks objects after everything else.
        if (cc[k].kind == AFF_CON_CONNECT && g.parent >= 0 && H.roundOf[g.parent] >= H.roundOf[bodyRef[cc[k].body_a]]) L.warpOk = 0;
        L.gcons.push_back(g);
      }
      kc.n_gcon = (int)L.gcons.size() - kc.first_gcon;
      if (kc.n_gcon > 32) { L.err = "more than 32 graph constraints in one candidate"; return AFF_ERR_CAPACITY; }
      kc.pose_slot = c.pose_body >= 0 ? objSlotOfBody[c.pose_body] : -1;
      for (int k = 0; k < 6; ++k) kc.pose_q6[k] = c.pose_q6[k];
    }
    return AFF_OK;
  };
  {
    unsigned nThreads = 1;
    if (const char* e = getenv("AFF_PLAN_THREADS")) nThreads = (unsigned)std::max(1, atoi(e));
    else if (nCands >= 8192) nThreads = std::min(4u, std::max(1u, std::thread::hardware_concurrency() / 4));
    nThreads = (unsigned)std::min<size_t>(nThreads, std::max<size_t>(1, nCands / 1024));
    std::vector<Local> locals(nThreads);
    std::vector<std::thread> pool;
    for (unsigned t = 0; t < nThreads; ++t) {
      const size_t lo = nCands * t / nThreads, hi = nCands * (t + 1) / nThreads;
      if (nThreads == 1) locals[0].rc = work(lo, hi, locals[0]);
      else pool.emplace_back([&, t, lo, hi]() { locals[t].rc = work(lo, hi, locals[t]); });
    }
    for (std::thread& th : pool) th.join();
    for (unsigned t = 0; t < nThreads; ++t) if (locals[t].rc != AFF_OK) { err = locals[t].err; return locals[t].rc; }
    for (unsigned t = 0; t < nThreads; ++t) {
      const size_t lo = nCands * t / nThreads, hi = nCands * (t + 1) / nThreads;
      const int bv = (int)H.vias.size(), ba = (int)H.acts.size(), bg = (int)H.gcons.size();
      for (size_t i = lo; i < hi; ++i) { H.cands[i].first_via += bv; H.cands[i].first_act += ba; H.cands[i].first_gcon += bg; }
      H.vias.insert(H.vias.end(), locals[t].vias.begin(), locals[t].vias.end());
      H.acts.insert(H.acts.end(), locals[t].acts.begin(), locals[t].acts.end());
      H.gcons.insert(H.gcons.end(), locals[t].gcons.begin(), locals[t].gcons.end());
      H.maxXdim = std::max(H.maxXdim, locals[t].maxXdim); H.maxSteps = std::max(H.maxSteps, locals[t].maxSteps);
      maxViaUnk = std::max(maxViaUnk, locals[t].maxViaUnk);
      if (!locals[t].warpOk) P.warp_ok = 0;
    }
  }

  // ---- scalars
  P.nJ = S.nJ; P.nDyn = (int)H.dyn.size(); P.nStatic = (int)H.stat.size(); P.nObj = nObj;
  {
    // static nodes by depth (parents come first in H.stat): the prologue computes a level at a time
    const int nS = (int)H.stat.size();
    std::vector<int> level(nS, 0);
    int nLev = nS ? 1 : 0;
    for (int n = 0; n < nS; ++n) {
      const int par = H.stat[n].parent;
      level[n] = (par == AFFK_WORLD) ? 0 : level[AFFK_STATIC_IDX(par)] + 1;
      nLev = std::max(nLev, level[n] + 1);
    }
    H.stat_level_off.assign((size_t)nLev + 1, 0);
    for (int n = 0; n < nS; ++n) H.stat_level_off[(size_t)level[n] + 1]++;
    for (int l = 0; l < nLev; ++l) H.stat_level_off[(size_t)l + 1] += H.stat_level_off[(size_t)l];
    H.stat_order.assign((size_t)std::max(nS, 1), 0);
    std::vector<int> fillAt(H.stat_level_off.begin(), H.stat_level_off.end() - (nLev ? 1 : 0));
    for (int n = 0; n < nS; ++n) H.stat_order[(size_t)fillAt[(size_t)level[n]]++] = n;
    P.nStatLevels = nLev;
  }
  P.nDynShapes = (int)H.dshape.size(); P.nDynBodies = (int)H.dbody.size();
  P.nStatShapes = (int)H.sshape.size(); P.nStatBodies = (int)H.sbody.size();
  P.nTrees = (int)S.bpTree.size(); P.nSS = (int)H.pairsSS.size(); P.nSB = (int)H.pairsSB.size();
  auto refOf = [&](int b) { return b >= 0 ? slotOfRef(bodyRef[b]) : -1; };
  for (const KNode& nd : H.dyn) { H.dyn_chain.push_back(nd.chain); H.dyn_under.push_back(nd.under_obj); }
  P.ns_base = refOf(S.ns_base);
  for (int k = 0; k < 2; ++k) { P.ns_upperarm[k] = refOf(S.ns_upperarm[k]); P.ns_forearm[k] = refOf(S.ns_forearm[k]); P.ns_hand[k] = refOf(S.ns_hand[k]); }
  {
    // Guard of the FK split of the block-per-candidate kernel: every frame the step loop reads (task effector /
    // reference / frame slots, null-space bodies, objects, the nodes objects get connected to) must come out of
    // the early rounds; if one does not, the step loop keeps all rounds (correct, only slower).
    const int nD = (int)H.dyn.size();
    auto late = [&](int slot) { return slot >= 0 && slot < nD && H.roundOf[slot] >= P.nRoundsCrit; };
    bool ok = true;
    for (const KTask& t : H.tasks) ok = ok && !late(t.eff) && !late(t.ref) && !late(t.frame);
    for (const KGraphCon& g : H.gcons) ok = ok && !late(g.parent_slot);
    for (int32_t on : H.obj_node) ok = ok && !late(on);
    for (int n = 0; n < nD; ++n) if (H.dyn[n].under_obj) ok = ok && !late(n);
    ok = ok && !late(P.ns_base);
    for (int k = 0; k < 2; ++k) ok = ok && !late(P.ns_upperarm[k]) && !late(P.ns_forearm[k]) && !late(P.ns_hand[k]);
    if (!ok) P.nRoundsCrit = P.nRounds;
  }
  P.rbj_static_chain = rbjStatic;
  P.bp_threshold = S.bpThreshold;
  P.nCands = (int)nCands;
  P.opts = opts;

  // ---- thread-per-candidate kernel: which node frames have to be kept in the node array.  Its FK walk carries
  // the frame of the node it has just computed in registers and writes the world records of the shapes on it at
  // once; the array is read again only for (a) parents that are not the node processed immediately before their
  // child, (b) frames of IK joints that can appear in the chain of a task / null-space body (Jacobian columns),
  // (c) task effector / reference / frame bodies, null-space bodies, objects and everything an object can be
  // connected to.  In a pizza `get` that leaves out the finger and fingertip frames: 12 of 31 nodes.
  {
    const int nD = (int)H.dyn.size();
    H.dyn_store.assign((size_t)std::max(nD, 1), 0);
    auto keepSlot = [&](int slot) { if (slot >= 0 && slot < nD) H.dyn_store[(size_t)slot] = 1; };
    // (a) processing order: nodes outside objects in index order, then objects and their subtrees
    std::vector<int> order;
    for (int pass = 0; pass < 2; ++pass) for (int n = 0; n < nD; ++n) if ((H.dyn[n].under_obj != 0) == (pass == 1)) order.push_back(n);
    for (size_t k = 0; k < order.size(); ++k) {
      const int n = order[k];
      if (H.dyn[n].kind == AFFK_NODE_OBJECT) continue;               // parents of objects: (c)
      const int ps = H.dyn_pslot[(size_t)n];
      if (k == 0 || order[k - 1] != ps) keepSlot(ps);
    }
    // (c) slots the step loop names, and the chains they can have
    uint32_t chains = 0u;
    std::vector<int> objParents;                                     // every slot an object can hang on
    for (int sl : H.obj_parent_slot) objParents.push_back(sl);
    for (const KGraphCon& g : H.gcons) if (g.kind == AFF_CON_CONNECT) objParents.push_back(g.parent_slot);
    uint32_t objParentChains = 0u;
    for (int sl : objParents) { keepSlot(sl); if (sl >= 0 && sl < nD) objParentChains |= H.dyn[(size_t)sl].chain; }
    // an object connected below another object inherits that object's parents' chains as well: already in the union
    auto useSlot = [&](int slot) {
      if (slot < 0) return;
      keepSlot(slot);
      if (slot < nD) { chains |= H.dyn[(size_t)slot].chain; if (H.dyn[(size_t)slot].under_obj) chains |= objParentChains; }
    };
    for (const KTask& k : H.tasks) { useSlot(k.eff); useSlot(k.ref); useSlot(k.frame); }
    useSlot(P.ns_base);
    for (int k = 0; k < 2; ++k) { useSlot(P.ns_upperarm[k]); useSlot(P.ns_forearm[k]); useSlot(P.ns_hand[k]); }
    for (int n : H.obj_node) keepSlot(n);
    for (int n = 0; n < nD; ++n) if (H.dyn[n].kind == AFFK_NODE_OBJECT) keepSlot(n);
    // (b) frames of the IK joints in those chains
    for (int cj = 0; cj < S.nJ; ++cj) if ((chains >> cj) & 1u) keepSlot(H.j_node[(size_t)cj]);
    P.sweep_joint_mask = chains;
  }

  // ---- thread-per-candidate kernel: candidates grouped by structure, so that the 32 rollouts of a warp
  // walk the same task list and switch / re-parent at the same steps (the values differ, the branches do not)
  {
    auto mix = [](uint64_t h, uint64_t v) { h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2); return h; };
    auto bitsOf = [](double d) { uint64_t u; std::memcpy(&u, &d, 8); return u; };
    std::vector<uint64_t> cls(nCands);
    for (size_t i = 0; i < nCands; ++i) {
      const KCand& kc = H.cands[i];
      uint64_t h = mix(mix((uint64_t)kc.first_task, (uint64_t)kc.n_tasks), (uint64_t)kc.n_steps);
      for (int d = 0; d <= kc.xdim; ++d) h = mix(h, (uint64_t)kc.via_off[d]);
      const int nv = kc.via_off[kc.xdim];
      for (int k = 0; k < nv; ++k) { const KVia& v = H.vias[kc.first_via + k]; h = mix(mix(h, bitsOf(v.t)), (uint64_t)v.flag); }
      for (int t = 0; t <= kc.n_tasks; ++t) h = mix(h, (uint64_t)kc.act_off[t]);
      const int na = kc.act_off[kc.n_tasks];
      for (int k = 0; k < na; ++k) { const KAct& a = H.acts[kc.first_act + k]; h = mix(mix(mix(h, bitsOf(a.t)), bitsOf(a.horizon)), (uint64_t)a.on); }
      for (int k = 0; k < kc.n_gcon; ++k) {
        const KGraphCon& g = H.gcons[kc.first_gcon + k];
        h = mix(mix(mix(mix(h, bitsOf(g.t)), (uint64_t)(g.kind * 4 + g.on)), (uint64_t)(g.slot * 64 + g.parent_slot)), (uint64_t)(g.parent_tree + 64));
      }
      cls[i] = mix(h, (uint64_t)(kc.pose_slot + 1));
    }
    H.sweep_order.resize(nCands);
    std::iota(H.sweep_order.begin(), H.sweep_order.end(), 0);
    bool sorted = true;
    for (size_t i = 1; i < nCands && sorted; ++i) sorted = cls[i] == cls[i - 1];
    if (!sorted) std::stable_sort(H.sweep_order.begin(), H.sweep_order.end(), [&](int32_t a, int32_t b) { return cls[a] < cls[b]; });
  }
  {
    int maxTasksS = 1;
    for (const KCand& kc : H.cands) maxTasksS = std::max(maxTasksS, kc.n_tasks);
    P.sweep_ok = (P.nDyn + (int)H.statpar_ref.size() <= AFFT_MAX_SLOTS && maxTasksS <= AFFT_MAX_TASKS && H.maxXdim <= AFFT_MAX_XDIM &&
                  P.nDynShapes <= AFFT_MAX_DSHAPE && P.nDynBodies <= AFFT_MAX_DBODY && P.nTrees <= AFFT_MAX_TREES && S.nJ <= AFFT_MAX_NJ) ? 1 : 0;
  }

  // FK table, second pass: operands that depend on the layout (see KFkEntry in rollout.cuh)
  auto patchFkAux = [&]() -> bool {
    const uint32_t spare = (uint32_t)(12 * (P.nDyn + P.nStatPar));
    for (KFkEntry& en : H.fktab) {
      const uint32_t op = (en.w0 >> 24) & 7u;
      uint32_t aux = (en.w1 >> 16) & 4095u;
      if (op == 0) en.w0 = (en.w0 & ~(4095u << 12)) | (spare << 12);                 // idle: store to the spare element
      else if (op == 1) aux += (uint32_t)P.o_node;                                    // COPY: parent element
      else if (op >= 3 && op <= 5 && !(en.w1 & 0x80u)) aux = (uint32_t)((op == 5) ? P.o_q : P.o_sn) + (en.w1 & 63u);
      if (aux > 4095u || spare > 4095u) return false;
      en.w1 = (en.w1 & 0xF000FFFFu) | (aux << 16);
    }
    return true;
  };

  // ---- shared-memory layout (doubles)
  int o = 0;
  auto take = [&](int n) { const int at = o; o += n; return at; };
  P.nStatPar = (int)H.statpar_ref.size();
  if (P.nDyn + P.nStatPar > 250) { err = "too many FK node slots"; return AFF_ERR_CAPACITY; }
  int maxTasks = 1;
  for (const KCand& kc : H.cands) maxTasks = std::max(maxTasks, kc.n_tasks);
  const int nSlots = std::max(P.nDyn + P.nStatPar, 1);
  P.via_unk = std::max(maxViaUnk, 1);
  P.via_stride = (P.via_unk * (P.via_unk + 1) + P.via_unk) | 1;   // matrix + solution, odd stride
  const int uTraj = P.via_stride * std::min(std::max(H.maxXdim, 1), AFFK_VIA_LANES);
  // live list (2*MAX_NEAR ints) + body trees (MAX_DBODY ints) + costly-pair scratch (32 ints + 32 doubles)
  const int collTail = AFFK_MAX_NEAR + AFFK_MAX_DBODY / 2 + 16 + 32;
  P.fixed_layout = (S.nJ <= AFFK_FIX_NJ && H.maxXdim <= AFFK_FIX_XDIM && maxTasks <= AFFK_FIX_TASKS && nSlots <= AFFK_FIX_SLOTS &&
                    P.nDynShapes <= AFFK_FIX_DSHAPE && P.nDynBodies <= AFFK_FIX_DBODY && P.nTrees <= AFFK_FIX_TREES &&
                    !getenv("AFF_TUNE_GENERIC_LAYOUT")) ? 1 : 0;
  if (P.fixed_layout) {
    // the compile-time layout of plan.h; only the node array (after the union) starts at a run-time offset
    P.o_q = AFFK_FIX_o_q; P.o_qdot = AFFK_FIX_o_qdot; P.o_dH = AFFK_FIX_o_dH; P.o_dq = AFFK_FIX_o_dq; P.o_sn = AFFK_FIX_o_sn; P.o_cs = AFFK_FIX_o_cs;
    P.o_dx = AFFK_FIX_o_dx; P.o_st = AFFK_FIX_o_st; P.o_xdes = AFFK_FIX_o_xdes; P.o_xcur = AFFK_FIX_o_xcur; P.o_geo = AFFK_FIX_o_geo;
    P.o_objrel = AFFK_FIX_o_objrel; P.o_objq6 = AFFK_FIX_o_objq6; P.o_chain = AFFK_FIX_o_chain; P.o_tdesc = AFFK_FIX_o_tdesc; P.o_rmask = AFFK_FIX_o_rmask;
    P.o_scratch = AFFK_FIX_o_scratch; P.o_J = AFFK_FIX_o_J; P.o_L = AFFK_FIX_o_L; P.o_rhs = AFFK_FIX_o_rhs;
    P.o_shp = AFFK_FIX_o_shp; P.o_baabb = AFFK_FIX_o_baabb; P.o_taabb = AFFK_FIX_o_taabb; P.o_live = AFFK_FIX_o_live;
    P.JS = AFFK_FIX_JS; P.LS = AFFK_FIX_LS;
    const int uIk = AFFK_FIX_IK_END - AFFK_FIX_o_scratch, uColl = AFFK_FIX_o_live + collTail - AFFK_FIX_o_scratch;
    P.o_node = AFFK_FIX_o_scratch + std::max(uTraj, std::max(uIk, uColl));
    P.warp_doubles = (P.o_node + 12 * nSlots + 1 + 1) & ~1;         // + the spare element idle FK lanes store to
    H.smemBytesPerWarp = sizeof(double) * (size_t)P.warp_doubles;
    if (!patchFkAux()) { err = "FK table: shared-memory index out of range"; return AFF_ERR_CAPACITY; }
    return AFF_OK;
  }
  P.o_node = take(12 * nSlots + 2);          // + the spare element idle FK lanes store to
  const int nJe = (S.nJ + 1) & ~1, xde = (std::max(H.maxXdim, 1) + 1) & ~1;
  P.o_q = take(nJe); P.o_qdot = take(nJe); P.o_dH = take(nJe); P.o_dq = take(nJe); P.o_sn = take(nJe); P.o_cs = take(nJe);
  P.o_dx = take(xde);
  P.o_st = take(3 * xde);
  P.o_xdes = take(xde); P.o_xcur = take(xde);
  P.o_geo = take(12 * maxTasks);
  P.o_objrel = take(12 * AFFK_MAX_OBJ); P.o_objq6 = take(6 * AFFK_MAX_OBJ);
  P.o_chain = take((P.nDyn + P.nStatPar + 1) / 2 + 1);      // uint32 per slot
  P.o_tdesc = take(4 * maxTasks);                            // 8 ints per task
  P.o_rmask = take(xde / 2);                                 // uint32 per J row
  // union region: trajectory solver scratch | J, L, rhs | collision scratch
  const int uBase = o;
  P.JS = (S.nJ | 1); P.LS = (H.maxXdim | 1);
  const int mRows = std::max(H.maxXdim, 1);
  const int uIk = mRows * P.JS + mRows * P.LS + mRows;
  const int uColl = 12 * std::max(P.nDynShapes, 1) + 6 * std::max(P.nDynBodies, 1) + 6 * std::max(P.nTrees, 1) + collTail;
  P.o_scratch = uBase;
  P.o_J = uBase; P.o_L = uBase + mRows * P.JS; P.o_rhs = P.o_L + mRows * P.LS;
  P.o_shp = uBase; P.o_baabb = P.o_shp + 12 * std::max(P.nDynShapes, 1); P.o_taabb = P.o_baabb + 6 * std::max(P.nDynBodies, 1);
  P.o_live = P.o_taabb + 6 * std::max(P.nTrees, 1);
  o = uBase + std::max(uTraj, std::max(uIk, uColl));
  P.warp_doubles = (o + 1) & ~1;
  H.smemBytesPerWarp = sizeof(double) * (size_t)P.warp_doubles;
  if (!patchFkAux()) { err = "FK table: shared-memory index out of range"; return AFF_ERR_CAPACITY; }
  return AFF_OK;
}

}   // namespace affk
