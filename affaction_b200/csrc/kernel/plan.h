// Device-side "plan": the flattened, index-only form of (scene, tasks,
// constraints, candidates) that the rollout kernel walks.  Built on the host by
// plan.cpp (structure only: no floating-point arithmetic happens there, so
// every rounded number the verdicts depend on is produced by the kernels).
//
// One warp runs one candidate rollout.  Bodies whose world pose can change
// during a rollout become "dynamic nodes" kept in shared memory; everything
// else is a "static node" whose world pose is computed once per batch by the
// prologue kernel and read from HBM/L2 through the read-only path.
#ifndef AFF_KERNEL_PLAN_H
#define AFF_KERNEL_PLAN_H

#include <stdint.h>

#include "affpredict.h"

#define AFFK_MAX_NJ 32        // IK joints: one lane each
#define AFFK_MAX_DYN 40       // dynamic nodes
#define AFFK_MAX_XDIM 32      // stacked task dimensions of one candidate
#define AFFK_MAX_TASKS 32     // tasks of one candidate
#define AFFK_MAX_M 16         // task rows active at the same time
#define AFFK_MAX_OBJ 2        // re-parentable objects per candidate
#define AFFK_MAX_DSHAPE 40    // shapes on dynamic nodes
#define AFFK_MAX_DBODY 32     // shape bodies on dynamic nodes
#define AFFK_MAX_NEAR 64      // body pairs found by the AABB test in one step
#define AFFK_MAX_PAIRS 512    // precompiled shape pairs per type list
#define AFFK_MAX_VIA_UNK 12   // capacity: unknown polynomial coefficients per 1-D window (KPlan::via_unk is the batch's bound)
#define AFFK_MAX_ROUNDS 40    // FK schedule rounds
#define AFFK_HORIZON 1.0       // trajectory horizon, src/ActionBase.cpp:137
#define AFFK_VIA_LANES 16     // trajectory dimensions solved at the same time (8 would fit a fifth block per SM, but two passes cost more)

// Caps of the per-thread state of the thread-per-candidate kernel (sweep.cuh); batches beyond them
// run on the warp-per-candidate kernel.
#ifndef AFFT_MAX_SLOTS
#define AFFT_MAX_SLOTS 48      // dynamic nodes + copies of the static frames the loop refers to
#define AFFT_MAX_TASKS 16
#define AFFT_MAX_XDIM 24
#define AFFT_MAX_DSHAPE 32
#define AFFT_MAX_DBODY 24
#define AFFT_MAX_TREES 6
#endif
#define AFFT_MAX_M AFFT_MAX_XDIM
#define AFFT_MAX_NJ AFFK_MAX_NJ
#define AFFT_MAX_COSTLY 32

// node reference: >= 0 dynamic node, -1 world, <= -2 static node (-2 - idx)
#define AFFK_WORLD (-1)
#define AFFK_STATIC_REF(i) (-2 - (i))
#define AFFK_STATIC_IDX(r) (-2 - (r))

enum { AFFK_NODE_FIXED = 0, AFFK_NODE_JOINT_IK = 1, AFFK_NODE_JOINT_CONST = 2, AFFK_NODE_OBJECT = 3 };

struct KNode
{
  int32_t parent;      // node reference
  int32_t kind;        // AFFK_NODE_*
  int32_t jtype;       // AFF_JNT_* for joint kinds
  int32_t qidx;        // JOINT_IK: jacobi index; OBJECT: object slot
  int32_t identityT;   // T is the identity: copy the parent
  int32_t under_obj;   // slot+1 if this node is a re-parentable object or below one
  int32_t tree;        // broad-phase tree the node belongs to by construction, -1 none
  uint32_t chain;      // IK joints (bit = jacobi index) that move this node by construction
  double qconst;       // JOINT_CONST value
  double T[12];        // A_JP (joints) or A_BP (fixed)
};

struct KShape
{
  int32_t type;        // AFF_SHAPE_*
  int32_t active0;     // RCSSHAPE_COMPUTE_DISTANCE at start
  int32_t node;        // dynamic shapes: shared-memory slot of the owning body; static shapes: node reference
  int32_t pad;
  double bound;        // radius of a sphere around the shape's centre that contains it (culling only)
  double ext[3];
  double A_CB[12];
};

// A body owning primitive shapes.
struct KShapeBody
{
  int32_t body;        // scene body id (result reporting, pair ordering)
  int32_t node;        // node reference
  int32_t first_shape, n_shapes;   // into the dynamic or static shape table
  int32_t explicit_bp; // listed as <Body> in the broad phase
  int32_t tree;        // static tree id or -1 (dynamic bodies not under an object)
  int32_t obj_slot;    // slot+1 if the body is a re-parentable object or below one (tree membership)
  int32_t toggle_slot; // slot+1 if the body IS the object (CollisionModelConstraint target)
  int32_t any_active0; // has an initially active shape
  int32_t pad;
};

// Precompiled shape pair of the collision model.  a/b: shape indices, >= 0 dynamic
// shape table, < 0 static shape table (-1 - idx).  Segment-segment list: `a` is
// dynamic; KP_SWAP says it is the SECOND argument in ascending-body-id order.
// Segment-box/cylinder list: `a` is the segment, `b` the box or cylinder.
struct __attribute__((aligned(16))) KPair { int32_t a, b, key, meta; };

// One lane's work in one FK round (see FKOP_* in rollout.cuh).
struct __attribute__((aligned(32))) KFkEntry { uint32_t w0, w1; double t0, t1, t2; };

struct KTask
{
  int32_t kind;
  int32_t eff, ref, frame;   // shared-memory slots, -1 = none / world (frame already resolved)
  int32_t axis;
  int32_t n_joints;
  int32_t jidx[AFF_MAX_TASK_JOINTS];  // jacobi index, or -1 - (static value slot) for constrained joints
  double jconst[AFF_MAX_TASK_JOINTS]; // value of constrained joints
  int32_t has_region;
  int32_t xoff;              // offset in the candidate's stacked x
  double region_min, region_max, region_dx_scaling;
};

struct KVia { double t, x, xd, xdd; int32_t flag; int32_t pad; };
struct KAct { double t, horizon; int32_t on; int32_t pad; };
// on: collision toggle direction; for a connect, 1 = the relative transform is the identity
// parent_tree: broad-phase tree of the new parent, or -2-s if the parent hangs below object s
struct KGraphCon { double t; int32_t kind; int32_t on; int32_t slot; int32_t parent; int32_t parent_slot; int32_t parent_tree; };

struct KCand
{
  int32_t first_task, n_tasks, xdim;
  int32_t first_via;               // vias sorted by (dimension, time)
  int32_t via_off[AFFK_MAX_XDIM + 1]; // per-dimension slice, relative to first_via
  int32_t first_act;               // activations sorted by (task, time)
  int32_t act_off[AFFK_MAX_TASKS + 1];
  int32_t first_gcon, n_gcon;
  int32_t n_steps;
  int32_t pose_slot;               // object slot whose q6 is overridden, -1 none
  double pose_q6[6];
  double end_abs, start_abs;
};

// Fixed shared-memory layout (doubles) for batches within these caps: every region except the node
// array starts at a compile-time offset and J / L have compile-time strides, so the kernel compiled
// with AFFK_FIXED_LAYOUT needs no address arithmetic per access.  buildPlan() writes the same numbers
// into KPlan, so the generic kernel (and the test emulator) run the identical layout.
#define AFFK_FIX_NJ 24         // IK joints
#define AFFK_FIX_XDIM 14       // task dimensions
#define AFFK_FIX_TASKS 8
#define AFFK_FIX_SLOTS 48      // node slots
#define AFFK_FIX_DSHAPE 20
#define AFFK_FIX_DBODY 16
#define AFFK_FIX_TREES 4
#define AFFK_FIX_JS 25
#define AFFK_FIX_LS 15
enum {
  AFFK_FIX_o_q = 0, AFFK_FIX_o_qdot = 24, AFFK_FIX_o_dH = 48, AFFK_FIX_o_dq = 72, AFFK_FIX_o_sn = 96, AFFK_FIX_o_cs = 120,
  AFFK_FIX_o_dx = 144, AFFK_FIX_o_st = 158, AFFK_FIX_o_xdes = 200, AFFK_FIX_o_xcur = 214,
  AFFK_FIX_o_geo = 228, AFFK_FIX_o_objrel = 324, AFFK_FIX_o_objq6 = 348, AFFK_FIX_o_chain = 360,
  AFFK_FIX_o_tdesc = 386, AFFK_FIX_o_rmask = 418,
  AFFK_FIX_o_scratch = 426,                       // union: trajectory solver | J, L, rhs | collision scratch
  AFFK_FIX_o_J = AFFK_FIX_o_scratch, AFFK_FIX_o_L = AFFK_FIX_o_J + AFFK_FIX_XDIM * AFFK_FIX_JS,
  AFFK_FIX_o_rhs = AFFK_FIX_o_L + AFFK_FIX_XDIM * AFFK_FIX_LS, AFFK_FIX_IK_END = AFFK_FIX_o_rhs + AFFK_FIX_XDIM,
  AFFK_FIX_o_shp = AFFK_FIX_o_scratch, AFFK_FIX_o_baabb = AFFK_FIX_o_shp + 12 * AFFK_FIX_DSHAPE,
  AFFK_FIX_o_taabb = AFFK_FIX_o_baabb + 6 * AFFK_FIX_DBODY, AFFK_FIX_o_live = AFFK_FIX_o_taabb + 6 * AFFK_FIX_TREES
};

// Everything the kernels need, passed by value.
struct KPlan
{
  int32_t nJ, nDyn, nStatic, nRounds, nObj;
  int32_t nRoundsCrit;             // FK rounds [0, nRoundsCrit) hold every node the step loop itself reads (plan_build.cpp: critB)
  int32_t nStatNear;               // static bodies [0, nStatNear) take part in the per-step AABB test (not explicit, active)
  uint32_t fk_obj_rounds[2], fk_const_rounds[2];   // bit r: FK round r holds an object node / a constant joint (rounds <= 64)
  int32_t nDynShapes, nDynBodies, nStatShapes, nStatBodies, nTrees;
  int32_t ns_base, ns_upperarm[2], ns_forearm[2], ns_hand[2];   // shared-memory slots, -1 = absent
  uint32_t rbj_static_chain;       // IK joints moving some non-object rigid-body-joint body
  double bp_threshold;

  // joints (jacobi order)
  const double* jq_init; const double* jq0; const double* jq_min; const double* jq_max;
  const double* jspeed; const double* jacc; const double* jwjl; const double* jwmetric;
  const int32_t* j_node;           // dynamic node of the joint frame
  const int32_t* j_type;

  const KNode* dyn;                // [nDyn], parents first
  const KNode* stat;               // [nStatic], parents first
  double* statA;                   // [nStatic*12] world transforms   (written by the prologue)
  double* statSC;                  // [nStatic*2]  sin/cos of constant joints
  double* dynSC;                   // [nDyn*2]     sin/cos of constant joints on dynamic nodes
  const KFkEntry* fktab;           // [nRounds*32] per-lane FK table
  const double* dynT;              // [nDyn*12] fixed transform of each dynamic node
  const double* dynQ;              // [nDyn] value of constant joints
  const int32_t* statpar_ref;      // [nStatPar] static nodes copied into shared memory as parents
  const int32_t* stat_order;       // [nStatic] static nodes sorted by depth below the world frame (prologue: one level at a time)
  const int32_t* stat_level_off;   // [nStatLevels + 1] level l = stat_order[stat_level_off[l] .. stat_level_off[l + 1])
  int32_t nStatLevels;
  const int32_t* obj_parent_slot;  // [nObj] initial parent slot of each object
  const int32_t* obj_parent_tree;  // [nObj] broad-phase tree of that parent
  const uint32_t* dyn_chain;       // [nDyn] IK joints moving each dynamic node by construction
  const int32_t* dyn_under;        // [nDyn] slot+1 if the node is a re-parentable object or below one
  int32_t nStatPar;
  // thread-per-candidate ("sweep") kernel, sweep.cuh
  const int32_t* dyn_pslot;        // [nDyn] slot of each dynamic node's parent (dynamic node or static-parent copy)
  const int32_t* node_shape_off;   // [nDyn + 1] CSR: dynamic shapes hanging on each dynamic node ...
  const int32_t* node_shape_idx;   // ... their indices in dshape (the sweep FK writes their world records)
  const int32_t* dyn_store;        // [nDyn] 1: the node's frame is read again after the FK walk (or by a non-adjacent child) and must be kept
  const int32_t* sweep_order;      // [nCands] candidates grouped by structure: thread i runs candidate sweep_order[i]
  uint32_t sweep_joint_mask;       // IK joints whose frames the sweep kernel keeps (they can appear in a task / null-space chain)
  int32_t sweep_ok;                // the batch fits the per-thread caps of sweep.cuh
  int32_t warp_ok;                 // the two-nodes-per-round FK schedule of rollout.cuh computes every possible parent of an object before the object

  const KShape* dshape; const KShapeBody* dbody;   // dynamic
  const KShape* sshape; const KShapeBody* sbody;   // static
  double* sshapeA;                 // [nStatShapes*12] world shape frames (prologue)
  double* sbodyAABB;               // [nStatBodies*6]                      (prologue)
  const KPair* pairsSS; const KPair* pairsSB;      // precompiled shape pairs by type
  int32_t nSS, nSB;
  double* sseg;                    // [nStatShapes*8] p(3), d(3), r, pad of static SSL / SPHERE shapes (prologue)
  const int32_t* obj_node;         // [nObj] dynamic node of each object slot
  const int32_t* obj_body;         // [nObj] scene body id
  const double* obj_q6;            // [nObj*6] scene values of the six rigid-body dofs

  const KTask* tasks; const KVia* vias; const KAct* acts; const KGraphCon* gcons; const KCand* cands;
  int32_t nCands;

  // shared-memory layout of one warp, in doubles from the warp base
  int32_t o_node, o_q, o_qdot, o_dH, o_dq, o_sn, o_cs, o_J, o_L, o_rhs, o_dx, o_st, o_xdes, o_xcur, o_geo,
          o_objrel, o_objq6, o_shp, o_baabb, o_taabb, o_scratch, o_live, o_chain, o_tdesc, o_rmask, warp_doubles;
  int32_t JS, LS;
  int32_t fixed_layout;            // 1: the offsets below are the AFFK_FIX_* constants (node array after the union)
  int32_t via_unk, via_stride;     // largest trajectory window of the batch (constrained values); solver scratch per lane (odd)                  // row strides of J and L (odd: conflict-free column access)

  aff_opts opts;
  aff_result* results;             // [nCands]
  double* qlog;                    // [nCands * (maxSteps+1) * nJ] or NULL
  int32_t qlog_rows;
};

#endif
