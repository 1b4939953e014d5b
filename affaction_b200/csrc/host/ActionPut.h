// "put <object> [<surface>] [frame <supportable frame>] [duration <s>]"
// Candidate generation and rollout description of the reference's ActionPut
// (src/ActionPut.cpp:91-303 candidates, :309-366 initialize, :368-476 tasks,
//  :478-597 trajectory).  One candidate per free <Supportable, Stackable> pair.
#ifndef AFF_HOST_ACTIONPUT_H
#define AFF_HOST_ACTIONPUT_H

#include "ActionBase.h"

namespace affgpu
{

class ActionPut : public ActionBase
{
public:
  ActionPut(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionPut(*this)); }

  bool initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank) override;
  size_t getNumSolutions() const override { return affordanceMap.size(); }
  std::vector<TaskSpec> createTasks() const override;
  ActivationSet createTrajectory(double t_start, double t_end) const override;
  using ActionBase::createTrajectory;
  std::vector<std::string> getManipulators() const override { return usedManipulators; }
  double getDurationHint() const override { return (putDown ? 0.6 : 1.0) * ActionBase::getDurationHint(); }
  std::string explain() const override { return explanation; }

  std::string objName, surfaceFrameName;

private:
  void init(const ActionScene& domain, const Graph& graph, const std::string& objectToPut, const std::string& surfaceToPutOn);

  std::vector<AffAffPair> affordanceMap;
  bool putDown, isObjCollidable, isPincerGrasped;
  double supportRegionX, supportRegionY;
  int polarAxisIdx;
  double downProjection[3];
  std::string whereOn, objBottomName, objGraspFrame, graspFrame, explanation;
  std::vector<std::string> fingerJoints, usedManipulators;
  std::vector<double> handOpen;
  std::string taskObjHandPos, taskHandSurfacePos, taskObjSurfacePosX, taskObjSurfacePosY, taskObjSurfacePosZ,
      taskObjSurfacePolar, taskHandPolar, taskHandObjPolar, taskSurfaceOri, taskFingers;
};

}   // namespace affgpu

#endif
