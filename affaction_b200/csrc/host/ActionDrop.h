// "drop <object> [<surface>]"
// Candidate generation and rollout description of the reference's ActionDrop
// (src/ActionDrop.cpp:51-159 candidates, :165-187 initialize, :189-214 tasks, :216-245 trajectory):
// one candidate per Supportable of the surface, closest first; 3 s; the object lands on the
// surface frame through a ConnectBodyConstraint with an identity transform.
#ifndef AFF_HOST_ACTIONDROP_H
#define AFF_HOST_ACTIONDROP_H

#include "ActionBase.h"

namespace affgpu
{

class ActionDrop : public ActionBase
{
public:
  ActionDrop(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionDrop(*this)); }

  bool initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank) override;
  size_t getNumSolutions() const override { return affordanceMap.size(); }
  std::vector<TaskSpec> createTasks() const override;
  ActivationSet createTrajectory(double t_start, double t_end) const override;
  using ActionBase::createTrajectory;
  std::vector<std::string> getManipulators() const override { return std::vector<std::string>(); }
  double getDurationHint() const override { return 3.0; }
  std::string explain() const override { return explanation; }

  std::string objectToDrop, surfaceName, graspFrameName;

private:
  std::vector<AffCapPair> affordanceMap;
  std::vector<std::string> fingerJoints;
  std::string taskInclination, taskPosition, taskFingers, explanation;
};

}   // namespace affgpu

#endif
