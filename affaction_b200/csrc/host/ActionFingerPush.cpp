#include "ActionFingerPush.h"

#include <cfloat>
#include <cmath>

namespace affgpu
{

static const double kFingersClosed = 1.32;   // reference src/ActionFingerPush.cpp:48

static std::vector<const Affordance*> affordancesOfType(const AffordanceEntity* e, Affordance::Type t)
{
  std::vector<const Affordance*> res;
  for (const Affordance& a : e->affordances) if (a.isA(t)) res.push_back(&a);
  return res;
}

ActionFingerPush::ActionFingerPush(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
{
  if (params.empty())
    throw ActionException("ActionFingerPush: needs at least 1 parameter, but received 0", ActionException::ParamInvalid);

  const AffordanceEntity* object = domain.getAffordanceEntity(params[0]);
  if (!object)
    throw ActionException("ERROR REASON: The " + params[0] + " is unknown. SUGGESTION: Use an object name that is defined in the environment",
                          ActionException::ParamNotFound);
  const auto pushables = affordancesOfType(object, Affordance::Type::PointPushable);
  if (pushables.empty())
    throw ActionException("ERROR REASON: The " + params[0] + " is not switchable. SUGGESTION: Replace this action with a more clever one.",
                          ActionException::ParamNotFound);
  if (pushables.size() != 1)
    throw ActionException("ERROR REASON: found " + std::to_string(pushables.size()) + " frames, but only 1 is supported",
                          ActionException::ParamInvalid);
  frameToBePoked = pushables[0]->frame;

  const Manipulator* theHand = nullptr;   // only set when a finger pokes: its finger joints are closed
  if (params.size() > 1) {
    const AffordanceEntity* pusher = domain.getAffordanceEntity(params[1]);
    if (!pusher)
      throw ActionException("ERROR REASON: The tool to push with " + params[1] +
                            " is unknown. SUGGESTION: Use an object name that is defined in the environment", ActionException::ParamNotFound);
    const auto pokables = affordancesOfType(pusher, Affordance::Type::PointPokable);
    if (pokables.empty())
      throw ActionException("ERROR REASON: The tool " + params[1] + " cannot be used to switch something on. SUGGESTION: Use another tool",
                            ActionException::ParamNotFound);
    if (pokables.size() != 1)
      throw ActionException("ActionFingerPush: found " + std::to_string(pokables.size()) + " pokables, but only 1 is supported",
                            ActionException::ParamInvalid);
    frameThatPokes = pokables[0]->frame;
  } else {
    const int pushBdy = graph.getBodyByName(frameToBePoked);
    if (pushBdy < 0) throw ActionException("ERROR REASON: frame " + frameToBePoked + " is not in the graph", ActionException::UnknownError);
    int closestFinger = -1;
    double dMin = DBL_MAX;
    for (const Manipulator* m : domain.getFreeManipulators(graph))      // occupied hands are not used
      for (const Capability& g : m->capabilities) {
        if (g.classType != Capability::Type::FingerpushCapability) continue;
        const int fbdy = graph.getBodyByName(g.frame);
        if (fbdy < 0) continue;
        double d2 = 0.0;
        for (int k = 0; k < 3; ++k) { const double d = graph.bodies[fbdy].A_BI.org[k] - graph.bodies[pushBdy].A_BI.org[k]; d2 += d * d; }
        const double d = std::sqrt(d2);
        if (d < dMin) { dMin = d; closestFinger = fbdy; theHand = m; }
      }
    if (closestFinger < 0)
      throw ActionException("ERROR REASON: Robot has no fingers to perform this action. SUGGESTION: Use a tool", ActionException::ParamNotFound);
    frameThatPokes = graph.bodies[closestFinger].name;
  }

  taskPushPos = "PushButton-XYZ-" + frameThatPokes + "-" + frameToBePoked;
  taskPushOri = "PushButton-Polar-" + frameThatPokes + "-" + frameToBePoked;
  taskFingers = frameThatPokes + "_fingers";
  if (theHand) fingerJoints = theHand->fingerJoints;

  explanation = "I'm pushing the " + params[0];
  if (params.size() > 1) explanation += " with the " + params[1];
}

std::vector<TaskSpec> ActionFingerPush::createTasks() const
{
  std::vector<TaskSpec> tasks;
  TaskSpec t;
  t.name = taskPushPos; t.controlVariable = "XYZ"; t.effector = frameThatPokes; t.refBdy = frameToBePoked;
  tasks.push_back(t);
  t = TaskSpec(); t.name = taskPushOri; t.controlVariable = "POLAR"; t.effector = frameThatPokes; t.refBdy = frameToBePoked;
  tasks.push_back(t);
  if (!fingerJoints.empty()) {
    t = TaskSpec(); t.name = taskFingers; t.controlVariable = "Joints"; t.jnts = fingerJoints;
    tasks.push_back(t);
  }
  return tasks;
}

// t_start .. t_prep (above the switch) .. t_touch .. t_down (pushed in) .. t_end (retracted)
ActivationSet ActionFingerPush::createTrajectory(double t_start, double t_end) const
{
  const double afterTime = 0.5;
  const double duration = t_end - t_start;
  const double t_prep = t_start + 0.5 * duration;
  const double t_touch = t_start + 0.7 * duration;
  const double t_down = t_start + 0.8 * duration;
  const double push_depth = -0.03, push_retract = 0.1;

  ActivationSet a1;
  a1.addActivation(t_start, true, 0.5, taskPushPos);
  a1.addActivation(t_start, true, 0.5, taskPushOri);
  a1.addActivation(t_end + afterTime, false, 0.5, taskPushPos);
  a1.addActivation(t_end, false, 0.5, taskPushOri);
  a1.add(t_prep, 0.0, 0.0, 0.0, 7, taskPushPos + " 0");
  a1.add(t_prep, 0.05, 0.0, 0.0, 7, taskPushPos + " 2");
  a1.addPositionConstraint(t_touch, 0.0, 0.0, 0.0, taskPushPos);
  a1.addPositionConstraint(t_down, 0.0, 0.0, push_depth, taskPushPos);
  a1.addPositionConstraint(t_end, 0.0, 0.0, push_retract, taskPushPos);
  a1.addPolarConstraint(t_touch, M_PI, 0.0, taskPushOri);
  a1.addPolarConstraint(t_end, M_PI, 0.0, taskPushOri);
  if (!fingerJoints.empty()) {
    a1.addActivation(t_start, true, 0.5, taskFingers);
    a1.addActivation(t_end + afterTime, false, 0.5, taskFingers);
    a1.addVectorConstraint(t_prep, std::vector<double>(fingerJoints.size(), kFingersClosed), taskFingers);
    a1.addVectorConstraint(t_end, std::vector<double>(fingerJoints.size(), kFingersClosed), taskFingers);
  }
  return a1;
}

}   // namespace affgpu
