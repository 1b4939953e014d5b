// ActionBase and the small task / constraint vocabulary the get / put / pour
// actions are written in.  Mirrors the reference's public surface for the
// predict step (src/ActionBase.h:131-147, src/ActionFactory.h:93-105): an
// action owns a list of solutions (Capability x Affordance candidates),
// initialize(rank) selects one, createTasks() and createTrajectory() describe
// its rollout, predict() forward-simulates it.
//
// Where the reference emits XML task strings for Rcs::TaskFactory and builds
// tropic::ActivationSet objects, this shim fills plain descriptor structs
// (include/affpredict.h) that cross the C ABI to the GPU.
#ifndef AFF_HOST_ACTIONBASE_H
#define AFF_HOST_ACTIONBASE_H

#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "ActionScene.h"
#include "TrajectoryPredictor.h"
#include "affpredict.h"

namespace affgpu
{

class ActionException : public std::runtime_error
{
public:
  enum ErrorCode { NoError = 0, ParamNotFound, ParamInvalid, KinematicallyImpossible, UnrecoverableError, UnknownError };
  ActionException(const std::string& msg, ErrorCode c) : std::runtime_error(msg), code(c) {}
  ErrorCode code;
};

// One <Task .../> of createTasksXML(), by name.
struct TaskSpec
{
  std::string name;
  std::string controlVariable;   // "XYZ" "X" "Y" "Z" "POLAR" "Inclination" "Joints"
  std::string effector, refBdy, refFrame;
  std::string axisDirection;     // "", "X", "Y", "Z"
  std::vector<std::string> jnts;
  bool hasRegion = false;
  double regionMin = 0.0, regionMax = 0.0, dxScaling = 0.0;
  int dim() const;
};

// The subset of tropic::ActivationSet the four actions use.
class ActivationSet
{
public:
  struct Entry
  {
    int kind;   // AFF_CON_*
    double t, x = 0, xd = 0, xdd = 0, horizon = 0;
    int flag = 0, dim = 0;
    std::string task, bodyA, bodyB;
  };
  void addActivation(double t, bool switchesOn, double horizon, const std::string& taskName);
  // add(t, x, xd, xdd, flag, "taskName dim")
  void add(double t, double x, double xd, double xdd, int flag, const std::string& trajNameND);
  void addPositionConstraint(double t, double x, double y, double z, const std::string& task, int flag = 7);   // tropic::PositionConstraint
  void addPolarConstraint(double t, double phi, double theta, const std::string& task, int flag = 7);          // tropic::PolarConstraint
  void addVectorConstraint(double t, const std::vector<double>& x, const std::string& task);     // tropic::VectorConstraint
  void addConnectBodyConstraint(double t, const std::string& child, const std::string& parent,
                                bool identityTransform = false);  // tropic::ConnectBodyConstraint (+ setConnectTransform(identity))
  void addCollisionModelConstraint(double t, const std::string& body, bool switchOn);            // tropic::CollisionModelConstraint
  void add(const ActivationSet& other);
  std::vector<Entry> entries;
};

// Flat descriptor tables for a set of candidates, ready for aff_predict_batch.
struct CandidateBatch
{
  std::vector<aff_task_desc> tasks;
  std::vector<aff_constraint> constraints;
  std::vector<aff_candidate> candidates;
  std::vector<std::vector<std::string>> taskNames;   // per candidate, for messages
  void append(const Graph& graph, const std::vector<TaskSpec>& tasks, const ActivationSet& tSet,
              int poseBody = -1, const double* poseQ6 = nullptr);
  // Same as append() for a candidate whose task list is that of candidate `like` of `proto` (sweeps: the tasks of
  // a command and hand do not change with the perturbation): the task records and names are copied, only the
  // trajectory is converted.
  void appendLike(const Graph& graph, const CandidateBatch& proto, size_t like, const ActivationSet& tSet,
                  int poseBody = -1, const double* poseQ6 = nullptr);

private:
  void appendConstraints(const Graph& graph, const std::vector<std::string>& names, const ActivationSet& tSet, aff_candidate& c);

public:
};

class ActionBase
{
public:
  ActionBase() : defaultDuration(10.0) {}
  virtual ~ActionBase() {}
  virtual std::unique_ptr<ActionBase> clone() const = 0;

  virtual ActivationSet createTrajectory(double t_start, double t_end) const = 0;
  ActivationSet createTrajectory() const { return createTrajectory(0.0, getDurationHint()); }
  virtual std::vector<TaskSpec> createTasks() const = 0;     // reference: createTasksXML()
  virtual std::vector<std::string> getManipulators() const = 0;
  virtual bool initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank) { return true; }
  virtual size_t getNumSolutions() const { return 1; }
  virtual double getDurationHint() const { return defaultDuration; }
  virtual std::string explain() const { return getActionCommand(); }

  // Forward-simulates the currently initialised solution on the GPU
  // (reference src/ActionBase.cpp:116-198).  Throws if the CUDA library fails.
  TrajectoryPredictor::PredictionResult predict(aff_scene* scene, const Graph& graph, double duration, double dt) const;

  // Appends the currently initialised solution to a batch: tasks + the
  // trajectory createTrajectory(5*dt, duration+5*dt) (src/ActionBase.cpp:136-139).
  void appendTo(CandidateBatch& batch, const Graph& graph, double duration, double dt) const;

  void setName(const std::string& n) { actionName = n; }
  std::string getName() const { return actionName; }
  void setActionParams(const std::vector<std::string>& p) { actionParams = p; }
  std::string getActionCommand() const;

protected:
  double defaultDuration;
  std::string actionName;
  std::vector<std::string> actionParams;
};

// Manipulator::getGraspingCapability: the grasp capability whose frame parents the object, or null.
const Capability* getGraspingCapability(const Graph& graph, const Manipulator* hand, int objBody);
// ActionBase::raycastSurface (reference src/ActionBase.cpp:271-313): the entity straight below the
// bottom of the object's bounding box, or null.
const AffordanceEntity* raycastSurface(const ActionScene& domain, const Graph& graph, int objBody);

// name -> constructor registry (reference src/ActionFactory.h:61-105).
class ActionFactory
{
public:
  // Splits `text` into words and builds the action; throws ActionException.
  static std::unique_ptr<ActionBase> create(const ActionScene& domain, const Graph& graph, const std::string& text);
};

}   // namespace affgpu

#endif
