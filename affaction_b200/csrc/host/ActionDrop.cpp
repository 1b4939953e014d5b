#include "ActionDrop.h"

#include <cmath>

namespace affgpu
{

static const double kFingersOpen = 0.01;   // reference src/ActionDrop.cpp:44

ActionDrop::ActionDrop(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
{
  if (params.empty()) throw ActionException("ActionDrop: No object given", ActionException::ParamNotFound);
  objectToDrop = params[0];

  const auto dropEntities = domain.getAffordanceEntities(objectToDrop);
  if (dropEntities.empty()) throw ActionException("Entity for " + objectToDrop + " not found in scene", ActionException::ParamNotFound);
  const Manipulator* graspingHand = nullptr;
  const AffordanceEntity* dropEntity = nullptr;
  for (const AffordanceEntity* e : dropEntities) if ((graspingHand = domain.getGraspingHand(graph, e))) { dropEntity = e; break; }
  if (!graspingHand)
    throw ActionException("ActionDrop: no hand is grasping the object " + objectToDrop, ActionException::KinematicallyImpossible);
  objectToDrop = dropEntity->bdyName;
  fingerJoints = graspingHand->fingerJoints;

  const int objBody = graph.getBodyByName(objectToDrop);
  const Capability* graspCapability = getGraspingCapability(graph, graspingHand, objBody);
  graspFrameName = graspCapability ? graspCapability->frame : std::string();
  taskInclination = graspingHand->name + "-InclinationY";
  taskPosition = graspingHand->name + "-XYZ";
  taskFingers = graspingHand->name + "_fingers";
  explanation = "I'm dropping the " + objectToDrop;

  // Only Supportables can be dropped on
  const AffordanceEntity* surface = nullptr;
  if (params.size() >= 2) surface = domain.getAffordanceEntity(params[1]);
  else {
    surface = raycastSurface(domain, graph, objBody);
    if (!surface)
      throw ActionException("ERROR: Can't drop the " + objectToDrop + " REASON: There is nothing under the " + objectToDrop +
                            " where I can put it on", ActionException::KinematicallyImpossible);
    surfaceName = surface->bdyName;
  }
  if (!surface) throw ActionException("ActionDrop: no surface entity to drop on found", ActionException::KinematicallyImpossible);

  for (const Affordance& a : surface->affordances)
    if (a.isA(Affordance::Type::Supportable)) affordanceMap.emplace_back(&a, graspCapability);
  if (affordanceMap.empty()) throw ActionException("ActionDrop: Cannot drop object on " + surface->name, ActionException::ParamNotFound);
  // The reference sorts by the distance to the grasp capability's frame and would dereference a null
  // capability here; without one the match order is kept.
  if (graspCapability) sort(graph, affordanceMap, 1.0, 0.0);

  if (!initialize(domain, graph, 0)) throw ActionException("Failed to initialize best solution", ActionException::ParamInvalid);
}

bool ActionDrop::initialize(const ActionScene&, const Graph&, size_t solutionRank)
{
  if (solutionRank >= getNumSolutions()) return false;
  surfaceName = std::get<0>(affordanceMap[solutionRank])->frame;
  return true;
}

std::vector<TaskSpec> ActionDrop::createTasks() const
{
  std::vector<TaskSpec> tasks;
  if (!graspFrameName.empty()) {
    // incline the hand (y axis of the grasp frame) so that the object does not roll over the fingers
    TaskSpec t;
    t.name = taskInclination; t.controlVariable = "Inclination"; t.effector = graspFrameName; t.axisDirection = "Y";
    tasks.push_back(t);
    t = TaskSpec(); t.name = taskPosition; t.controlVariable = "XYZ"; t.effector = graspFrameName;
    tasks.push_back(t);
  }
  TaskSpec f;
  f.name = taskFingers; f.controlVariable = "Joints"; f.jnts = fingerJoints;
  tasks.push_back(f);
  return tasks;
}

ActivationSet ActionDrop::createTrajectory(double t_start, double t_end) const
{
  const double duration = t_end - t_start;
  ActivationSet a1;
  if (!graspFrameName.empty()) {
    a1.addActivation(t_start, true, 0.5, taskPosition);      // the hand stays where it is
    a1.addActivation(t_end, false, 0.5, taskPosition);
    a1.addActivation(t_start, true, 0.5, taskInclination);
    a1.addActivation(t_end, false, 0.5, taskInclination);
    a1.add(t_start + 0.5 * duration, M_PI_2, 0.0, 0.0, 7, taskInclination + " 0");
  }
  a1.addActivation(t_start, true, 0.5, taskFingers);
  a1.addActivation(t_end, false, 0.5, taskFingers);
  a1.addVectorConstraint(t_end, std::vector<double>(fingerJoints.size(), kFingersOpen), taskFingers);
  // the dropped object becomes a child of the surface frame, placed onto it
  a1.addConnectBodyConstraint(t_end, objectToDrop, surfaceName, true);
  return a1;
}

}   // namespace affgpu
