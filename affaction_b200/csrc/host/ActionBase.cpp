#include "ActionBase.h"

#include <algorithm>
#include <cstring>
#include <sstream>

namespace affgpu
{

int TaskSpec::dim() const
{
  if (controlVariable == "XYZ") return 3;
  if (controlVariable == "POLAR") return 2;
  if (controlVariable == "Joints" || controlVariable == "Joint") return (int)jnts.size();
  return 1;
}

// ------------------------------------------------------------ ActivationSet
void ActivationSet::addActivation(double t, bool on, double horizon, const std::string& task)
{
  Entry e{};
  e.kind = AFF_CON_ACTIVATION; e.t = t; e.flag = on ? 1 : 0; e.horizon = horizon; e.task = task;
  entries.push_back(e);
}

void ActivationSet::add(double t, double x, double xd, double xdd, int flag, const std::string& trajNameND)
{
  const size_t sp = trajNameND.rfind(' ');
  if (sp == std::string::npos) throw ActionException("via point needs \"task dim\": " + trajNameND, ActionException::ParamInvalid);
  Entry e{};
  e.kind = AFF_CON_VIA; e.t = t; e.x = x; e.xd = xd; e.xdd = xdd; e.flag = flag;
  e.task = trajNameND.substr(0, sp); e.dim = std::stoi(trajNameND.substr(sp + 1));
  entries.push_back(e);
}

void ActivationSet::addPositionConstraint(double t, double x, double y, double z, const std::string& task, int flag)
{
  const double v[3] = {x, y, z};
  for (int i = 0; i < 3; ++i) add(t, v[i], 0.0, 0.0, flag, task + " " + std::to_string(i));
}

void ActivationSet::addPolarConstraint(double t, double phi, double theta, const std::string& task, int flag)
{
  add(t, phi, 0.0, 0.0, flag, task + " 0");
  add(t, theta, 0.0, 0.0, flag, task + " 1");
}

void ActivationSet::addVectorConstraint(double t, const std::vector<double>& x, const std::string& task)
{
  for (size_t i = 0; i < x.size(); ++i) add(t, x[i], 0.0, 0.0, 7, task + " " + std::to_string(i));
}

void ActivationSet::addConnectBodyConstraint(double t, const std::string& child, const std::string& parent, bool identityTransform)
{
  Entry e{};
  e.kind = AFF_CON_CONNECT; e.t = t; e.bodyA = child; e.bodyB = parent; e.flag = 1; e.dim = identityTransform ? 1 : 0;
  entries.push_back(e);
}

void ActivationSet::addCollisionModelConstraint(double t, const std::string& body, bool on)
{
  Entry e{};
  e.kind = AFF_CON_COLLISION; e.t = t; e.bodyA = body; e.flag = on ? 1 : 0;
  entries.push_back(e);
}

void ActivationSet::add(const ActivationSet& other)
{
  entries.insert(entries.end(), other.entries.begin(), other.entries.end());
}

// ----------------------------------------------------------- CandidateBatch
static int bodyId(const Graph& g, const std::string& n, const char* what)
{
  if (n.empty()) return -1;
  const int b = g.getBodyByName(n);
  if (b < 0) throw ActionException(std::string("ERROR: unknown body '") + n + "' in " + what, ActionException::ParamNotFound);
  return b;
}

static int axisIndex(const std::string& a)
{
  if (a == "X") return 0;
  if (a == "Y") return 1;
  return 2;
}

void CandidateBatch::append(const Graph& graph, const std::vector<TaskSpec>& specs, const ActivationSet& tSet,
                            int poseBody, const double* poseQ6)
{
  aff_candidate c;
  std::memset(&c, 0, sizeof(c));
  c.first_task = (int32_t)tasks.size();
  c.first_constraint = (int32_t)constraints.size();
  c.pose_body = poseBody;
  if (poseBody >= 0 && poseQ6) for (int k = 0; k < 6; ++k) c.pose_q6[k] = poseQ6[k];
  std::vector<std::string> names;
  for (const TaskSpec& s : specs) {
    // addTasks(): a task whose name already exists is not added twice (src/ActionBase.cpp:95-110)
    if (std::find(names.begin(), names.end(), s.name) != names.end()) continue;
    aff_task_desc t;
    std::memset(&t, 0, sizeof(t));
    const std::string& cv = s.controlVariable;
    if (cv == "XYZ") t.kind = AFF_TASK_XYZ;
    else if (cv == "X") t.kind = AFF_TASK_X;
    else if (cv == "Y") t.kind = AFF_TASK_Y;
    else if (cv == "Z") t.kind = AFF_TASK_Z;
    else if (cv == "POLAR") t.kind = AFF_TASK_POLAR;
    else if (cv == "Inclination") t.kind = AFF_TASK_INCLINATION;
    else if (cv == "Joints" || cv == "Joint") t.kind = AFF_TASK_JOINTS;
    else throw ActionException("ERROR: controlVariable '" + cv + "' is not supported by the GPU predictor", ActionException::ParamInvalid);
    t.effector = bodyId(graph, s.effector, s.name.c_str());
    t.ref_body = bodyId(graph, s.refBdy, s.name.c_str());
    t.ref_frame = bodyId(graph, s.refFrame, s.name.c_str());
    t.axis = axisIndex(s.axisDirection);
    if (t.kind == AFF_TASK_JOINTS) {
      if (s.jnts.size() > AFF_MAX_TASK_JOINTS) throw ActionException("ERROR: too many joints in task " + s.name, ActionException::ParamInvalid);
      t.n_joints = (int32_t)s.jnts.size();
      for (size_t i = 0; i < s.jnts.size(); ++i) {
        t.joints[i] = graph.getJointByName(s.jnts[i]);
        if (t.joints[i] < 0) throw ActionException("ERROR: unknown joint '" + s.jnts[i] + "'", ActionException::ParamNotFound);
      }
    } else if (t.effector < 0) {
      throw ActionException("ERROR: task " + s.name + " has no effector", ActionException::ParamInvalid);
    }
    t.has_region = s.hasRegion ? 1 : 0;
    t.region_min = s.regionMin; t.region_max = s.regionMax; t.region_dx_scaling = s.dxScaling;
    tasks.push_back(t);
    names.push_back(s.name);
  }
  c.n_tasks = (int32_t)names.size();
  appendConstraints(graph, names, tSet, c);
  candidates.push_back(c);
  taskNames.push_back(names);
}

void CandidateBatch::appendLike(const Graph& graph, const CandidateBatch& proto, size_t like, const ActivationSet& tSet,
                                int poseBody, const double* poseQ6)
{
  const aff_candidate& pc = proto.candidates[like];
  aff_candidate c;
  std::memset(&c, 0, sizeof(c));
  c.first_task = (int32_t)tasks.size();
  c.first_constraint = (int32_t)constraints.size();
  c.pose_body = poseBody;
  if (poseBody >= 0 && poseQ6) for (int k = 0; k < 6; ++k) c.pose_q6[k] = poseQ6[k];
  tasks.insert(tasks.end(), proto.tasks.begin() + pc.first_task, proto.tasks.begin() + pc.first_task + pc.n_tasks);
  c.n_tasks = pc.n_tasks;
  const std::vector<std::string>& names = proto.taskNames[like];
  appendConstraints(graph, names, tSet, c);
  candidates.push_back(c);
  taskNames.push_back(names);
}

void CandidateBatch::appendConstraints(const Graph& graph, const std::vector<std::string>& names, const ActivationSet& tSet, aff_candidate& c)
{
  for (const ActivationSet::Entry& e : tSet.entries) {
    aff_constraint k;
    std::memset(&k, 0, sizeof(k));
    k.kind = e.kind; k.t = e.t; k.x = e.x; k.xd = e.xd; k.xdd = e.xdd; k.horizon = e.horizon; k.flag = e.flag; k.dim = e.dim;
    k.task = -1; k.body_a = -1; k.body_b = -1;
    if (e.kind == AFF_CON_VIA || e.kind == AFF_CON_ACTIVATION) {
      const auto it = std::find(names.begin(), names.end(), e.task);
      // tropic ignores constraints on trajectories that do not exist; so do we
      if (it == names.end()) continue;
      k.task = (int32_t)(it - names.begin());
    } else {
      k.body_a = bodyId(graph, e.bodyA, "constraint");
      if (e.kind == AFF_CON_CONNECT) k.body_b = bodyId(graph, e.bodyB, "constraint");
    }
    constraints.push_back(k);
  }
  c.n_constraints = (int32_t)constraints.size() - c.first_constraint;
}

// --------------------------------------------------------------- ActionBase
std::string ActionBase::getActionCommand() const
{
  std::string cmd = actionName;
  for (const auto& p : actionParams) { cmd += " "; cmd += p; }
  return cmd;
}

const Capability* getGraspingCapability(const Graph& graph, const Manipulator* hand, int objBody)
{
  for (const Capability& c : hand->capabilities)
    if (c.isGrasp() && graph.getBodyByName(c.frame) == graph.bodies[objBody].parent) return &c;
  return nullptr;
}

const AffordanceEntity* raycastSurface(const ActionScene& domain, const Graph& graph, int objBody)
{
  double mn[3], mx[3], dMin = 0.0;
  if (!graph.computeBodyAABB(objBody, mn, mx)) return nullptr;
  const double from[3] = {0.5 * (mn[0] + mx[0]), 0.5 * (mn[1] + mx[1]), mn[2] - 1.0e-8}, dir[3] = {0.0, 0.0, -1.0};
  const int below = graph.closestRigidBodyInDirection(from, dir, &dMin);
  return below >= 0 ? domain.getAffordanceEntity(graph.bodies[below].name) : nullptr;
}

void ActionBase::appendTo(CandidateBatch& batch, const Graph& graph, double duration, double dt) const
{
  const double delay = 5.0 * dt;                       // src/ActionBase.cpp:138
  batch.append(graph, createTasks(), createTrajectory(delay, duration + delay));
}

TrajectoryPredictor::PredictionResult ActionBase::predict(aff_scene* scene, const Graph& graph, double duration, double dt) const
{
  CandidateBatch batch;
  appendTo(batch, graph, duration, dt);
  return TrajectoryPredictor::predictBatch(scene, graph, batch, dt)[0];
}

}   // namespace affgpu
