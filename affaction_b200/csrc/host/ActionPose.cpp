#include "ActionPose.h"

namespace affgpu
{

ActionPose::ActionPose(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
{
  if (params.empty())
    throw ActionException("ERROR REASON: Could not find pose DEVELOPER: no model state name given", ActionException::ParamNotFound);
  const std::string& mdlStateName = params[0];
  const int timeStamp = params.size() > 1 ? std::stoi(params[1]) : 0;
  const ModelState* state = domain.getModelState(mdlStateName, timeStamp);
  if (!state)
    throw ActionException("ERROR REASON: Could not find pose DEVELOPER: Model state with name not found: " + mdlStateName,
                          ActionException::ParamNotFound);
  // graph order, unconstrained joints the state names (RCSGRAPH_FOREACH_JOINT, src/ActionPose.cpp:79-92)
  for (const Joint& j : graph.joints) {
    if (j.constrained) continue;
    for (const auto& js : state->joints)
      if (js.first == j.name) { jnts.emplace_back(j.name, j.q_init, js.second); break; }
  }
  for (const Manipulator& m : domain.manipulators) usedManipulators.push_back(m.name);
  explanation = "Moving to pose " + mdlStateName;
}

std::vector<TaskSpec> ActionPose::createTasks() const
{
  std::vector<TaskSpec> tasks;
  for (const auto& ji : jnts) {
    TaskSpec t;
    t.name = std::get<0>(ji); t.controlVariable = "Joint"; t.jnts = {std::get<0>(ji)};
    tasks.push_back(t);
  }
  return tasks;
}

ActivationSet ActionPose::createTrajectory(double t_start, double t_end) const
{
  ActivationSet a1;
  for (const auto& ji : jnts) {
    const std::string& jntName = std::get<0>(ji);
    a1.addActivation(t_start, true, 0.5, jntName);
    a1.addActivation(t_end + 0.5, false, 0.5, jntName);
    a1.add(t_end, std::get<2>(ji), 0.0, 0.0, 7, jntName + " 0");
  }
  return a1;
}

}   // namespace affgpu
