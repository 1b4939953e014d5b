// integration/rcs_graph_adapter.h — fills an aff_scene_desc (include/affpredict.h) from a live RcsGraph.
//
// THIS HEADER IS FOR THE REFERENCE TREE.  It includes Rcs headers, which this repository does not have
// (Rcs is an un-vendored dependency of HRI-EU/AffAction, CMakeLists.txt:96-103), so it is NOT compiled or
// tested here; everything else in this repository gets its scene from a flattened .affscene file instead
// (tools/flatten_scene.py).  It uses only the public RcsGraph / RcsBody / RcsJoint / RcsShape fields and
// macros the reference itself touches on the predict path:
//   RCSGRAPH_FOREACH_BODY, BODY->id, ->parentId, ->firstChildId / ->lastChildId, ->rigid_body_joints,
//   ->A_BI, ->name                                   src/TrajectoryPredictor.cpp:113-118,648-664; src/ActionGet.cpp:441-458
//   RCSGRAPH_TRAVERSE_JOINTS, JNT->jointIndex, ->constrained, ->q_min / ->q_max, ->name
//                                                     src/TrajectoryPredictor.cpp:810-818
//   shape->computeType & RCSSHAPE_COMPUTE_DISTANCE, shape->type (RCSSHAPE_MESH, RCSSHAPE_REFFRAME)
//                                                     src/CollisionModelConstraint.h:118-140
//   graph->q, graph->nJ, graph->dof                   src/TrajectoryPredictor.cpp:582-585,733
//   RcsBroadPhase (trees, bodies, distanceThreshold)  src/ActionBase.cpp:127-131
//
// Use (ActionComponent, next to where it stores `domain`, src/ActionComponent.cpp:78):
//   affgpu::RcsSceneAdapter adapter;
//   const aff_scene_desc* desc = adapter.fill(graph, broadphase);
//   aff_scene_create(desc, /*device=*/0, &gpuScene);          // once per loaded graph
//   ... per action: adapter.updateState(graph) + aff_scene_create again, or keep one aff_scene per state
#ifndef AFF_INTEGRATION_RCS_GRAPH_ADAPTER_H
#define AFF_INTEGRATION_RCS_GRAPH_ADAPTER_H

#include <Rcs_graph.h>
#include <Rcs_body.h>
#include <Rcs_joint.h>
#include <Rcs_shape.h>
#include <Rcs_broadphase.h>
#include <Rcs_typedef.h>

#include <cfloat>
#include <cstring>
#include <string>
#include <vector>

#include "affpredict.h"

namespace affgpu
{

class RcsSceneAdapter
{
public:
  // Bodies in graph order (RcsGraph stores parents before children: body id = array index), joints in
  // the order of each body's joint chain (jntId -> nextId), shapes with a distance or toggleable role.
  const aff_scene_desc* fill(const RcsGraph* graph, const RcsBroadPhase* bp)
  {
    clear();
    std::vector<int> bodyIndexOfId(graph->nBodies, -1);
    RCSGRAPH_FOREACH_BODY(graph)
    {
      if (BODY->id == -1) continue;                          // deleted slot, as in src/TrajectoryPredictor.cpp:650
      bodyIndexOfId[BODY->id] = (int)parent_.size();
      parent_.push_back(BODY->parentId >= 0 ? bodyIndexOfId[BODY->parentId] : -1);
      rbj_.push_back(BODY->rigid_body_joints ? 1 : 0);
      push12(abp_, &BODY->A_BP);
      names_.push_back(BODY->name);
      // joints of the body, parent side first
      firstJoint_.push_back((int32_t)jtype_.size());
      int nj = 0;
      for (const RcsJoint* JNT = RCSJOINT_BY_ID(graph, BODY->jntId); JNT; JNT = RCSJOINT_BY_ID(graph, JNT->nextId), ++nj) {
        jtype_.push_back(jointType(JNT->type));
        jcon_.push_back(JNT->constrained ? 1 : 0);
        push12(ajp_, &JNT->A_JP);
        qinit_.push_back(graph->q->ele[JNT->jointIndex]);   // the LIVE state is the start state of a prediction
        q0_.push_back(JNT->q0); qmin_.push_back(JNT->q_min); qmax_.push_back(JNT->q_max);
        speed_.push_back(JNT->speedLimit); acc_.push_back(JNT->accLimit);
        wjl_.push_back(JNT->weightJL); wmetric_.push_back(JNT->weightMetric);
        jointIndex_.push_back(JNT->jointIndex);
        jointNames_.push_back(JNT->name);
      }
      nJoints_.push_back(nj);
      // shapes: primitives that take part in distance computation now, or can be switched on by a
      // CollisionModelConstraint (every non-mesh, non-frame shape: src/CollisionModelConstraint.h:118-140)
      firstShape_.push_back((int32_t)stype_.size());
      int ns = 0;
      for (unsigned int i = 0; i < BODY->nShapes; ++i) {
        const RcsShape* SHAPE = &BODY->shapes[i];
        const int t = shapeType(SHAPE->type);
        if (t < 0) continue;
        stype_.push_back(t);
        sdist_.push_back((SHAPE->computeType & RCSSHAPE_COMPUTE_DISTANCE) ? 1 : 0);
        push12(acb_, &SHAPE->A_CB);
        // extents as the kernels read them: SSL / SPHERE {radius, -, length}, BOX full widths, CYLINDER {radius, -, height}
        ext_.push_back(SHAPE->extents[0]); ext_.push_back(SHAPE->extents[1]); ext_.push_back(SHAPE->extents[2]);
        ++ns;
      }
      nShapes_.push_back(ns);
    }
    // broad phase: tree roots, explicit bodies, threshold (pizza: g_scenario_pizza.xml:264-271)
    for (unsigned int i = 0; i < bp->nTrees; ++i) trees_.push_back(bodyIndexOfId[bp->trees[i].bodyId]);
    for (unsigned int i = 0; i < bp->nBodies; ++i) if (!bp->bodies[i].hasSubTree) bpBodies_.push_back(bodyIndexOfId[bp->bodies[i].bodyId]);
    desc_.bp_threshold = bp->distanceThreshold;

    desc_.n_bodies = (int32_t)parent_.size(); desc_.n_joints = (int32_t)jtype_.size(); desc_.n_shapes = (int32_t)stype_.size();
    desc_.body_parent = parent_.data(); desc_.body_first_joint = firstJoint_.data(); desc_.body_n_joints = nJoints_.data();
    desc_.body_first_shape = firstShape_.data(); desc_.body_n_shapes = nShapes_.data(); desc_.body_rbj = rbj_.data();
    desc_.body_A_BP = abp_.data();
    desc_.joint_type = jtype_.data(); desc_.joint_constrained = jcon_.data(); desc_.joint_A_JP = ajp_.data();
    desc_.joint_q_init = qinit_.data(); desc_.joint_q0 = q0_.data(); desc_.joint_q_min = qmin_.data(); desc_.joint_q_max = qmax_.data();
    desc_.joint_speed_limit = speed_.data(); desc_.joint_acc_limit = acc_.data();
    desc_.joint_weight_jl = wjl_.data(); desc_.joint_weight_metric = wmetric_.data();
    desc_.shape_type = stype_.data(); desc_.shape_distance = sdist_.data(); desc_.shape_A_CB = acb_.data(); desc_.shape_extents = ext_.data();
    desc_.n_bp_trees = (int32_t)trees_.size(); desc_.bp_tree_root = trees_.data();
    desc_.n_bp_bodies = (int32_t)bpBodies_.size(); desc_.bp_body = bpBodies_.data();
    // bodies of the elbow / wrist null-space terms (src/TrajectoryPredictor.cpp:459-463,492-493,540-542)
    desc_.ns_base = indexOf("base_footprint");
    desc_.ns_upperarm[0] = indexOf("upperarm_left"); desc_.ns_upperarm[1] = indexOf("upperarm_right");
    desc_.ns_forearm[0] = indexOf("forearm_left"); desc_.ns_forearm[1] = indexOf("forearm_right");
    desc_.ns_hand[0] = indexOf("hand_left"); desc_.ns_hand[1] = indexOf("hand_right");
    return &desc_;
  }

  // New joint values only (same topology): cheaper than fill() between two predictions.
  void updateState(const RcsGraph* graph)
  {
    for (size_t j = 0; j < jointIndex_.size(); ++j) qinit_[j] = graph->q->ele[jointIndex_[j]];
  }

  const aff_scene_desc* desc() const { return &desc_; }
  const std::vector<std::string>& bodyNames() const { return names_; }
  const std::vector<std::string>& jointNames() const { return jointNames_; }

  int indexOf(const char* name) const
  {
    for (size_t i = 0; i < names_.size(); ++i) if (names_[i] == name) return (int)i;
    return -1;
  }
  const std::string& bodyName(int index) const { return names_[(size_t)index]; }

private:
  static void push12(std::vector<double>& v, const HTr* A)
  {
    for (int k = 0; k < 3; ++k) v.push_back(A->org[k]);
    for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) v.push_back(A->rot[i][k]);
  }
  static int jointType(int rcsType)
  {
    switch (rcsType) {
      case RCSJOINT_TRANS_X: return AFF_JNT_TRANS_X; case RCSJOINT_TRANS_Y: return AFF_JNT_TRANS_Y; case RCSJOINT_TRANS_Z: return AFF_JNT_TRANS_Z;
      case RCSJOINT_ROT_X: return AFF_JNT_ROT_X; case RCSJOINT_ROT_Y: return AFF_JNT_ROT_Y; default: return AFF_JNT_ROT_Z;
    }
  }
  static int shapeType(int rcsType)      // -1: no distance function on the predict path (MESH, REFFRAME, ...)
  {
    switch (rcsType) {
      case RCSSHAPE_SSL: return AFF_SHAPE_SSL; case RCSSHAPE_SPHERE: return AFF_SHAPE_SPHERE;
      case RCSSHAPE_BOX: return AFF_SHAPE_BOX; case RCSSHAPE_CYLINDER: return AFF_SHAPE_CYLINDER; default: return -1;
    }
  }
  void clear()
  {
    parent_.clear(); firstJoint_.clear(); nJoints_.clear(); firstShape_.clear(); nShapes_.clear(); rbj_.clear(); abp_.clear();
    jtype_.clear(); jcon_.clear(); ajp_.clear(); qinit_.clear(); q0_.clear(); qmin_.clear(); qmax_.clear(); speed_.clear();
    acc_.clear(); wjl_.clear(); wmetric_.clear(); jointIndex_.clear(); stype_.clear(); sdist_.clear(); acb_.clear(); ext_.clear();
    trees_.clear(); bpBodies_.clear(); names_.clear(); jointNames_.clear();
    std::memset(&desc_, 0, sizeof(desc_));
  }

  aff_scene_desc desc_;
  std::vector<int32_t> parent_, firstJoint_, nJoints_, firstShape_, nShapes_, jtype_, stype_, trees_, bpBodies_;
  std::vector<uint8_t> rbj_, jcon_, sdist_;
  std::vector<double> abp_, ajp_, qinit_, q0_, qmin_, qmax_, speed_, acc_, wjl_, wmetric_, acb_, ext_;
  std::vector<unsigned int> jointIndex_;
  std::vector<std::string> names_, jointNames_;
};

}   // namespace affgpu

#endif
