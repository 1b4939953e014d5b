// "pose <model state> [<time stamp>]"
// The reference's ActionPose (src/ActionPose.cpp:54-144): one "Joint" task per unconstrained joint
// named in the model state, driven to the state's value at t_end.  One candidate.
#ifndef AFF_HOST_ACTIONPOSE_H
#define AFF_HOST_ACTIONPOSE_H

#include <tuple>

#include "ActionBase.h"

namespace affgpu
{

class ActionPose : public ActionBase
{
public:
  ActionPose(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionPose(*this)); }

  std::vector<TaskSpec> createTasks() const override;
  ActivationSet createTrajectory(double t_start, double t_end) const override;
  using ActionBase::createTrajectory;
  std::vector<std::string> getManipulators() const override { return usedManipulators; }
  std::string explain() const override { return explanation; }

private:
  std::vector<std::tuple<std::string, double, double>> jnts;   // name, q_curr, q_des
  std::vector<std::string> usedManipulators;
  std::string explanation;
};

}   // namespace affgpu

#endif
