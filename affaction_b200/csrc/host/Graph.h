// Host-side kinematic graph: the part of RcsGraph / RcsBroadPhase that the
// predict path reads.  Loaded from the flattened `.affscene` text format that
// tools/flatten_scene.py writes from the reference's XML scene files.
//
// The host graph is used for candidate generation only (name lookup, the
// distance heuristic of sort(), lift-height ray cast); all rollout arithmetic
// happens on the GPU from the aff_scene_desc exported by toDesc().
#ifndef AFF_HOST_GRAPH_H
#define AFF_HOST_GRAPH_H

#include <array>
#include <string>
#include <unordered_map>
#include <vector>

#include "affpredict.h"

namespace affgpu
{

struct HTr
{
  double org[3];
  double rot[3][3];
  static HTr identity();
  static HTr from12(const double* v);
  void to12(double* v) const;
};

HTr compose(const HTr& child, const HTr& parent);

struct Joint
{
  std::string name;
  int type = 0;
  bool constrained = false;
  double q_init = 0, q0 = 0, q_min = 0, q_max = 0, speedLimit = 0, accLimit = 0, weightJL = 1, weightMetric = 1;
  HTr A_JP;
  int body = -1;
  int jacobiIndex = -1;
};

struct Shape
{
  int type = 0;
  bool distance = false;
  double extents[3] = {0, 0, 0};
  HTr A_CB;
  int body = -1;
};

struct Body
{
  std::string name;
  int parent = -1;
  bool rigid_body_joints = false;
  HTr A_BP;
  int firstJoint = 0, nJoints = 0, firstShape = 0, nShapes = 0;
  HTr A_BI;   // filled by computeForwardKinematics()
};

class Graph
{
public:
  // Throws std::runtime_error on malformed input.
  static Graph load(const std::string& affsceneFile, std::vector<std::string>* tail = nullptr);
  // The inverse of toDesc(): a Graph from a plain-C scene description plus the names the description does
  // not carry.  This is how the shim gets its graph in the REFERENCE tree, where the description comes from
  // the live RcsGraph (integration/rcs_graph_adapter.h) instead of an .affscene file.
  static Graph fromDesc(const aff_scene_desc& d, const std::vector<std::string>& bodyNames,
                        const std::vector<std::string>& jointNames, const std::string& graphName = "");

  int getBodyByName(const std::string& name) const;        // -1 if absent
  int getBodyByNameNoCase(const std::string& name) const;  // RcsGraph_getBodyByNameNoCase
  int getJointByName(const std::string& name) const;
  bool isChild(int body, int ancestor) const;               // RcsBody_isChild (strict descendant)
  int numDistanceShapes(int body) const;                    // RcsBody_numDistanceShapes
  std::vector<int> children(int body) const;

  // RcsGraph_setState on the initial state.
  void computeForwardKinematics();
  // Re-parent `child` under `parent` keeping its world pose: what a fired
  // ConnectBodyConstraint leaves behind (used to chain get -> put on the host).
  void connectBody(int child, int parent, bool identityTransform = false);
  // State a finished rollout leaves behind (what the live graph looks like after the action ran):
  // every ConnectBody / CollisionModel constraint is applied at the step it fired, with the joint
  // state of that step (qlog row), then the joints take the final row.  qlog: rows x nJ, jacobi order.
  void applyRollout(const struct aff_constraint* cons, int nCons, const double* qlog, int rows, double dt);
  // After re-parenting, a body may sit behind its new parent's subtree in the arrays; the plain-C scene
  // description (and the plan compiler behind it) lists parents before children.  Stable re-sort of the
  // body array (relative order kept), all body indices remapped.  Body ids handed out before the call are
  // stale afterwards (names are not).  Returns true if anything moved.
  bool restoreParentFirstOrder();
  // World transforms of EVERY body for every row of a rollout's joint log: what the reference keeps in
  // PredictionResult::bodyTransforms ((nSteps+1) x nBodies x 12: org + rot row-major per body,
  // src/TrajectoryPredictor.cpp:111-118,356-363) and ActionComponent replays as the ghost animation
  // (src/ActionComponent.cpp:429-462).  Replayed on a copy of the graph from the GPU's q(t): row 0 is the
  // initial state, row i+1 the state after step i with the constraints that fired in it applied.
  // out: rows x bodies.size() x 12 doubles.  *this is not modified.
  void rolloutTransforms(const struct aff_constraint* cons, int nCons, const double* qlog, int rows, double dt, double* out) const;

  // RcsGraph_computeBodyAABB over shapes with the distance flag. false if none.
  bool computeBodyAABB(int body, double xyzMin[3], double xyzMax[3]) const;
  // RcsBody_closestRigidBodyInDirection: first body hit by a ray, -1 if none.
  int closestRigidBodyInDirection(const double pt[3], const double dir[3], double* dMin) const;

  // Plain-C view for the C ABI / oracle; pointers stay valid while *this lives
  // and is not modified.
  const aff_scene_desc* toDesc();

  std::string name;
  std::vector<Body> bodies;
  std::vector<Joint> joints;
  std::vector<Shape> shapes;
  double bpThreshold = 0.0;
  std::vector<int> bpTrees, bpBodies;
  int nJ = 0;

private:
  std::unordered_map<std::string, int> bodyIndex_, jointIndex_;
  // storage behind toDesc()
  aff_scene_desc desc_{};
  std::vector<int32_t> d_parent_, d_fj_, d_nj_, d_fs_, d_ns_, d_jtype_, d_stype_, d_tree_, d_bpb_;
  std::vector<uint8_t> d_rbj_, d_jcon_, d_sdist_;
  std::vector<double> d_abp_, d_ajp_, d_qinit_, d_q0_, d_qmin_, d_qmax_, d_speed_, d_acc_, d_wjl_, d_wm_, d_acb_, d_ext_;
};

}   // namespace affgpu

#endif
