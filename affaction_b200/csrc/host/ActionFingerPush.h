// "finger_push <object to poke> [<object that pokes>]"  (also registered as "poke")
// The reference's ActionFingerPush (src/ActionFingerPush.cpp:67-226 construction, :228-250 tasks,
// :266-306 trajectory): the closest free FingerpushCapability, or a held PointPokable tool, is moved
// onto the object's PointPushable frame with opposing z axes.  One candidate, 16 s.
#ifndef AFF_HOST_ACTIONFINGERPUSH_H
#define AFF_HOST_ACTIONFINGERPUSH_H

#include "ActionBase.h"

namespace affgpu
{

class ActionFingerPush : public ActionBase
{
public:
  ActionFingerPush(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionFingerPush(*this)); }

  std::vector<TaskSpec> createTasks() const override;
  ActivationSet createTrajectory(double t_start, double t_end) const override;
  using ActionBase::createTrajectory;
  std::vector<std::string> getManipulators() const override { return usedManipulators; }
  double getDurationHint() const override { return 2.0 * 8.0; }
  std::string explain() const override { return explanation; }

  std::string frameToBePoked, frameThatPokes;

private:
  std::vector<std::string> fingerJoints, usedManipulators;
  std::string taskPushPos, taskPushOri, taskFingers, explanation;
};

}   // namespace affgpu

#endif
