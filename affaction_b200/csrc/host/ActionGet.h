// "get <object> [<hand>] [from <x>] [liftHeight <h>] [duration <s>]"
// Candidate generation and rollout description of the reference's ActionGet
// (src/ActionGet.cpp:79-500 candidates, :506-810 trajectories, :839-942 tasks).
#ifndef AFF_HOST_ACTIONGET_H
#define AFF_HOST_ACTIONGET_H

#include "ActionBase.h"

namespace affgpu
{

class ActionGet : public ActionBase
{
public:
  enum class GraspType { Other = 0, PowerGrasp, BallGrasp, TopGrasp, RimGrasp, OrientedGrasp };

  ActionGet(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionGet(*this)); }

  bool initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank) override;
  size_t getNumSolutions() const override { return affordanceMap.size(); }
  std::vector<TaskSpec> createTasks() const override;
  ActivationSet createTrajectory(double t_start, double t_end) const override;
  using ActionBase::createTrajectory;
  std::vector<std::string> getManipulators() const override { return usedManipulators; }
  std::string explain() const override { return "I'm getting the " + objectName; }

  static const char* graspTypeToString(GraspType g);

  // Knobs the synthetic sweep perturbs per rollout.
  double liftHeight, preGraspDist;
  GraspType graspType;
  std::string objectName, capabilityFrame, affordanceFrame;

private:
  void init(const ActionScene& domain, const Graph& graph, const std::string& objectToGet,
            const std::string& manipulator, const std::string& whereFrom);
  ActivationSet trajectoryOriented(double t_start, double t_grasp, double t_end) const;
  ActivationSet trajectoryBallGrasp(double t_start, double t_grasp, double t_end) const;
  ActivationSet trajectoryTopGrasp(double t_start, double t_grasp, double t_end) const;

  std::vector<AffCapPair> affordanceMap;
  double shoulderBase;
  bool handOver, isObjCollidable;
  std::string whereFrom, handOverHand;
  std::vector<std::string> fingerJoints, usedManipulators;
  std::vector<double> handOpen, handClosed;
  std::string taskObjHandPos, taskObjPosX, taskObjPosY, taskObjPosZ, taskObjOri, taskHandObjOri, taskFingers, taskHandOverHand;
};

}   // namespace affgpu

#endif
