// Host-side plan compiler: (scene description, task table, constraint table,
// candidates) -> index-only tables for the rollout kernel (plan.h).
// Pure C++ (no CUDA) so the test-only warp emulator can use it as well.
#ifndef AFF_KERNEL_PLAN_BUILD_H
#define AFF_KERNEL_PLAN_BUILD_H

#include <string>
#include <vector>

#include "plan.h"

namespace affk
{

// Deep copy of an aff_scene_desc plus derived index tables.
struct SceneModel
{
  int nB = 0, nQ = 0, nS = 0, nJ = 0;
  std::vector<int32_t> parent, firstJoint, nJoints, firstShape, nShapes, jtype, stype, bpTree, bpBody, jacobi, ikJoint, jointBody;
  std::vector<uint8_t> rbj, jcon, sdist, explicitBp;
  std::vector<double> A_BP, A_JP, qInit, q0, qMin, qMax, speed, acc, wjl, wmetric, A_CB, ext;
  double bpThreshold = 0.0;
  int ns_base = -1, ns_upperarm[2] = {-1, -1}, ns_forearm[2] = {-1, -1}, ns_hand[2] = {-1, -1};

  // returns "" or an error message
  std::string init(const aff_scene_desc* d);
};

// All tables of one batch, host side.  Pointers in `plan` are left null; the
// caller (abi.cu or the emulator) points them at device / host copies.
struct HostPlan
{
  KPlan plan{};
  std::vector<double> jq_init, jq0, jq_min, jq_max, jspeed, jacc, jwjl, jwmetric, obj_q6;
  std::vector<KFkEntry> fktab;
  std::vector<int32_t> j_node, j_type, rounds, statpar_ref, obj_parent_slot, obj_parent_tree, dyn_under, obj_node, obj_body, dyn_pslot, sweep_order, roundOf, node_shape_off, node_shape_idx, dyn_store, stat_order, stat_level_off;
  std::vector<uint32_t> dyn_chain;
  std::vector<double> dynT, dynQ;
  std::vector<KPair> pairsSS, pairsSB;
  std::vector<KNode> dyn, stat;
  std::vector<KShape> dshape, sshape;
  std::vector<KShapeBody> dbody, sbody;
  std::vector<KTask> tasks;
  std::vector<KVia> vias;
  std::vector<KAct> acts;
  std::vector<KGraphCon> gcons;
  std::vector<KCand> cands;
  int maxSteps = 0, maxXdim = 0;
  size_t smemBytesPerWarp = 0;
};

// Returns AFF_OK or a negative aff_status; `err` receives the message.
int buildPlan(const SceneModel& scene, const aff_task_desc* tasks, size_t nTasks, const aff_constraint* cons, size_t nCons,
              const aff_candidate* cands, size_t nCands, const aff_opts& opts, HostPlan& out, std::string& err);

}   // namespace affk

#endif
