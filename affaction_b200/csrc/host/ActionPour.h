// "pour <from> <into> [<litres>] [duration <s>]"
// Rollout description of the reference's ActionPour (src/ActionPour.cpp:57-91 parameters,
// :93-241 init, :247-279 tasks, :303-364 trajectory).  One candidate; default duration 20 s.
#ifndef AFF_HOST_ACTIONPOUR_H
#define AFF_HOST_ACTIONPOUR_H

#include "ActionBase.h"

namespace affgpu
{

class ActionPour : public ActionBase
{
public:
  ActionPour(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionPour(*this)); }

  std::vector<TaskSpec> createTasks() const override;
  ActivationSet createTrajectory(double t_start, double t_end) const override;
  using ActionBase::createTrajectory;
  std::vector<std::string> getManipulators() const override { return usedManipulators; }
  std::string explain() const override { return explanation; }

  std::string bottle, glas, roboBaseFrame;   // frames of the two openings, base frame
  bool glasInHand, bottleInRightHand;
  double pouringVolume;

private:
  void init(const ActionScene& domain, const Graph& graph, const std::string& objectToPourFrom, const std::string& objToPourInto,
            double amountToPour, const std::string& roboBase);

  std::string taskRelPos, taskBottleOri, taskGlasOri, taskGlasPosX, taskGlasPosZ, explanation;
  std::vector<std::string> usedManipulators, particlesToPour;
};

}   // namespace affgpu

#endif
