// Second call site of the predict step: the pre-execution re-check of a trajectory on the
// runtime controller (reference src/TrajectoryComponent.cpp:222-230 onSimulateTrajectory,
// :285-340 onCheckAndSetTrajectory, :346-399 checkerThread).  Same kernel, batch of one, with
// the predicted joint array kept for the ghost animation.
//
// The reference publishes events on its entity; here the events of one call are returned as
// values (CheckOutcome) for the caller's event loop to publish.
#ifndef AFF_HOST_TRAJECTORYCOMPONENT_H
#define AFF_HOST_TRAJECTORYCOMPONENT_H

#include <mutex>

#include "ActionBase.h"

namespace affgpu
{

class TrajectoryComponent
{
public:
  struct CheckOutcome
  {
    bool checked = false;              // a prediction ran
    bool publishActionResult = false;  // "ActionResult"(actionOk, quality, message)
    bool actionOk = false;
    double quality = 0.0;
    std::string message;
    bool publishSetTrajectory = false; // "SetTrajectory"(tSet): the trajectory goes to the controller
    TrajectoryPredictor::PredictionResult result;
  };

  TrajectoryComponent(aff_scene* scene, const Graph& graph, double dt);
  // the runtime controller's task list (the predictor works on a copy of it)
  void setControllerTasks(std::vector<TaskSpec> controllerTasks) { tasks = std::move(controllerTasks); }

  // tSet = the trajectory to check against the controller's tasks
  CheckOutcome onSimulateTrajectory(const ActivationSet& tSet) { return checkerThread(describe(tSet), true); }
  CheckOutcome onCheckAndSetTrajectory(const ActivationSet& tSet) { return onCheckAndSetTrajectory(describe(tSet)); }
  // the same with tasks + trajectory already resolved into descriptors (one candidate)
  CheckOutcome onSimulateTrajectory(const CandidateBatch& one) { return checkerThread(one, true); }
  CheckOutcome onCheckAndSetTrajectory(const CandidateBatch& one);
  CheckOutcome checkerThread(const CandidateBatch& one, bool simulateOnly);

  void onEmergencyStop() { eStop = true; }
  void onEmergencyRecover() { eStop = false; }
  void onEnableTrajectoryCheck(bool enable) { enableTrajectoryCheck = enable; }

  // predictor->getPredictionArray(): q(t) of the last checked trajectory, rows x nJ (IK joints in
  // jacobi order), row 0 = start state.  Read by the ghost animation.
  const std::vector<double>& getPredictionArray(int* rows, int* cols) const { *rows = tPredRows; *cols = tPredCols; return tPred; }

private:
  CandidateBatch describe(const ActivationSet& tSet) const { CandidateBatch b; b.append(graph, tasks, tSet); return b; }

  aff_scene* scene;
  const Graph& graph;
  std::vector<TaskSpec> tasks;
  double dt;
  bool eStop, enableTrajectoryCheck;
  std::mutex checkerThreadMtx;        // reentrancy lock of checkerThread
  std::vector<double> tPred;
  int tPredRows, tPredCols;
};

}   // namespace affgpu

#endif
