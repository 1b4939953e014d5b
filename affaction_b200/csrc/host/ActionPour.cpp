#include "ActionPour.h"

#include <algorithm>
#include <cmath>

namespace affgpu
{

static std::vector<const Affordance*> affordancesOfType(const AffordanceEntity* e, Affordance::Type t)
{
  std::vector<const Affordance*> res;
  for (const Affordance& a : e->affordances) if (a.isA(t)) res.push_back(&a);
  return res;
}

static double attributeOr(const Affordance* a, const char* key, double fallback)
{
  auto it = a->attributes.find(key);
  return it == a->attributes.end() ? fallback : std::stod(it->second);
}

ActionPour::ActionPour(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
  : glasInHand(false), bottleInRightHand(true), pouringVolume(0.1)
{
  if (params.size() < 2)
    throw ActionException("ERROR REASON: Action has less than two parameters. At least two parameters are needed:"
                          "The container to pour from, and the container to pour into. SUGGESTION: Correct action command.",
                          ActionException::ParamInvalid);
  defaultDuration = 20.0;
  auto it = std::find(params.begin(), params.end(), "duration");
  if (it != params.end() && it + 1 != params.end()) {
    defaultDuration = std::stod(*(it + 1));
    params.erase(it, it + 2);
  }
  double amountToPour = pouringVolume;
  if (params.size() > 2) amountToPour = std::stod(params[2]);
  if (amountToPour < 0.0)
    throw ActionException("ERROR REASON: The amount to pour has a negative value. It must be equal or larger than zero."
                          "SUGGESTION: Pour only liquid amounts with positive volume.", ActionException::ParamInvalid);
  init(domain, graph, params[0], params[1], amountToPour, "base_footprint");
}

void ActionPour::init(const ActionScene& domain, const Graph& graph, const std::string& objectToPourFrom, const std::string& objToPourInto,
                      double amountToPour, const std::string& roboBase)
{
  pouringVolume = amountToPour;
  const AffordanceEntity* pourFromAff = domain.getAffordanceEntity(objectToPourFrom);
  if (!pourFromAff)
    throw ActionException("ERROR REASON: The " + objectToPourFrom + " to pour from is unknown. SUGGESTION: Use an object name that is defined in the environment",
                          ActionException::ParamNotFound);
  const auto openings = affordancesOfType(pourFromAff, Affordance::Type::Pourable);
  if (openings.empty())
    throw ActionException("ERROR REASON: The " + objectToPourFrom + " is not a container that can be poured from."
                          "SUGGESTION: Choose another object to pour from.", ActionException::ParamNotFound);
  const AffordanceEntity* pourToAff = domain.getAffordanceEntity(objToPourInto);
  if (!pourToAff)
    throw ActionException("ERROR REASON: The " + objToPourInto + " to pour into is unknown. SUGGESTION: Use an object name that is defined in the environment",
                          ActionException::ParamNotFound);
  const auto containers = affordancesOfType(pourToAff, Affordance::Type::Containable);
  if (containers.empty())
    throw ActionException("ERROR REASON: The " + objToPourInto + " to pour into is not a container that can be poured into. "
                          "SUGGESTION: Choose another object to pour into.", ActionException::ParamNotFound);
  const Manipulator* pouringHand = domain.getGraspingHand(graph, pourFromAff);
  if (!pouringHand)
    throw ActionException("ERROR REASON: The " + objectToPourFrom + " to pour from is not held in a hand."
                          " SUGGESTION: First get the object " + objectToPourFrom + " before performing this action", ActionException::ParamNotFound);
  usedManipulators.push_back(pouringHand->name);
  if (openings.size() != 1 || containers.size() != 1)
    throw ActionException("ERROR REASON: Can't deal with more than 1 opening or container", ActionException::ParamInvalid);
  const int bottleBdy = graph.getBodyByName(openings[0]->frame), glasBdy = graph.getBodyByName(containers[0]->frame), baseBdy = graph.getBodyByName(roboBase);
  if (bottleBdy < 0 || glasBdy < 0 || baseBdy < 0)
    throw ActionException("ERROR REASON: pouring frames are missing in the graph", ActionException::UnknownError);
  bottle = graph.bodies[bottleBdy].name;
  glas = graph.bodies[glasBdy].name;
  roboBaseFrame = graph.bodies[baseBdy].name;

  // performLiquidTransition (src/ActionPour.cpp:377-461): only its overflow test decides feasibility.
  // The flattened scene carries max_volume but not the liquid ingredients; current volumes count as
  // what the attributes say ("volume", default 0 for the target and "enough" for the source).
  {
    const auto fromContainers = affordancesOfType(pourFromAff, Affordance::Type::Containable);
    if (fromContainers.size() != 1)
      throw ActionException("ERROR REASON: The " + pourFromAff->name + " that is poured from has " + std::to_string(fromContainers.size()) +
                            " containers - currently only one is supported", ActionException::ParamInvalid);
    const double fromVolume = attributeOr(fromContainers[0], "volume", pouringVolume);
    const double toVolume = attributeOr(containers[0], "volume", 0.0), toMax = attributeOr(containers[0], "max_volume", 0.0);
    const double volumeToBePoured = std::max(0.0, std::min(pouringVolume, fromVolume));
    if (volumeToBePoured + toVolume > toMax)
      throw ActionException("ERROR REASON: The container " + containers[0]->frame + " does only hold " + std::to_string(toMax) +
                            " liters, pouring would overflow it. SUGGESTION: Pour less than " + std::to_string(toMax - toVolume) + " liters",
                            ActionException::KinematicallyImpossible);
  }

  // entities standing in / on the source move over to the target at t_up (src/ActionPour.cpp:176-180)
  {
    const int fromBody = graph.getBodyByName(pourFromAff->bdyName);
    std::vector<int> stack = graph.children(fromBody);
    while (!stack.empty()) {
      const int b = stack.back();
      stack.pop_back();
      for (const AffordanceEntity& e : domain.entities) if (e.bdyName == graph.bodies[b].name) particlesToPour.push_back(e.bdyName);
      for (int c : graph.children(b)) stack.push_back(c);
    }
    if (!particlesToPour.empty())
      throw ActionException("ERROR REASON: The " + objectToPourFrom + " holds solid items; moving them to the target container "
                            "(ConnectBodyConstraint with an explicit transform) is not part of the GPU predict path", ActionException::UnknownError);
  }

  taskRelPos = bottle + "-" + glas + "-" + roboBase + "-XYZ";
  taskBottleOri = bottle + "-POLAR";
  taskGlasOri = glas + "-POLAR";
  taskGlasPosX = glas + "-X";
  // taskGlasPosZ keeps the empty name the reference leaves it with (src/ActionPour.cpp:183-187)

  // Left or right: compare the y components in the robot base frame, whose y axis points left (:189-216)
  const HTr& A_RI = graph.bodies[baseBdy].A_BI;
  auto inBase = [&](const double* p, double* out) {
    const double d[3] = {p[0] - A_RI.org[0], p[1] - A_RI.org[1], p[2] - A_RI.org[2]};
    for (int i = 0; i < 3; ++i) out[i] = A_RI.rot[i][0] * d[0] + A_RI.rot[i][1] * d[1] + A_RI.rot[i][2] * d[2];
  };
  double rGlas[3], rBottle[3];
  inBase(graph.bodies[glasBdy].A_BI.org, rGlas);
  inBase(graph.bodies[bottleBdy].A_BI.org, rBottle);
  bottleInRightHand = rBottle[1] < rGlas[1];

  const Manipulator* receivingHand = domain.getGraspingHand(graph, pourToAff);
  if (receivingHand) { usedManipulators.push_back(pouringHand->name); glasInHand = true; }
  else glasInHand = false;

  explanation = "I'm pouring " + std::to_string(pouringVolume) + " liters from the " + pourFromAff->name + " into the " + pourToAff->name;
}

std::vector<TaskSpec> ActionPour::createTasks() const
{
  std::vector<TaskSpec> tasks;
  TaskSpec t;
  t.name = taskRelPos; t.controlVariable = "XYZ"; t.effector = bottle; t.refBdy = glas; t.refFrame = roboBaseFrame;
  tasks.push_back(t);
  t = TaskSpec(); t.name = taskBottleOri; t.controlVariable = "POLAR"; t.effector = bottle;
  tasks.push_back(t);
  t = TaskSpec(); t.name = taskGlasOri; t.controlVariable = "POLAR"; t.effector = glas;
  tasks.push_back(t);
  t = TaskSpec(); t.name = taskGlasPosX; t.controlVariable = "X"; t.effector = glas; t.refFrame = roboBaseFrame;
  tasks.push_back(t);
  t = TaskSpec(); t.name = taskGlasPosZ; t.controlVariable = "Z"; t.effector = glas; t.refFrame = roboBaseFrame;
  tasks.push_back(t);
  return tasks;
}

// t_start .. t_prep (tip above the rim, 30 deg) .. t_up (aligned, 150 deg) .. t_down .. t_end (apart, 10 deg)
ActivationSet ActionPour::createTrajectory(double t_start, double t_end) const
{
  const double duration = t_end - t_start;
  const double t_prep = t_start + 0.15 * duration;
  const double t_up = t_start + 0.45 * duration;
  const double t_down = t_start + 0.9 * duration;
  const double afterTime = 0.5;
  const double thetaTilt = bottleInRightHand ? M_PI_2 : -M_PI_2;
  const double d_separate = bottleInRightHand ? -0.25 : +0.25;
  const double heightAboveGlas = 0.08;
  const double deg = M_PI / 180.0;

  ActivationSet a1;
  a1.addActivation(t_start, true, 0.5, taskRelPos);
  a1.addActivation(t_end + afterTime, false, 0.5, taskRelPos);
  a1.add(t_prep, 0.6 * d_separate, 0.0, 0.0, 7, taskRelPos + " 1");
  a1.add(t_prep + 0.5 * (t_up - t_prep), 0.0, 0.0, 0.0, 7, taskRelPos + " 1");
  a1.add(t_prep, heightAboveGlas, 0.0, 0.0, 7, taskRelPos + " 2");
  a1.addPositionConstraint(t_up, 0.0, 0.0, 0.0, taskRelPos);
  a1.addPositionConstraint(t_up + 0.5 * (t_down - t_up), 0.0, 0.0, heightAboveGlas, taskRelPos, 1);
  a1.addPositionConstraint(t_down, 0.0, d_separate, heightAboveGlas, taskRelPos, 1);
  a1.addPositionConstraint(t_end, 0.0, d_separate, heightAboveGlas, taskRelPos);

  a1.addActivation(t_start, true, 0.5, taskBottleOri);
  a1.addActivation(t_end + afterTime, false, 0.5, taskBottleOri);
  a1.addPolarConstraint(t_prep, 30.0 * deg, thetaTilt, taskBottleOri, 1);
  a1.addPolarConstraint(t_up, 150.0 * deg, thetaTilt, taskBottleOri);
  a1.addPolarConstraint(t_end, 10.0 * deg, thetaTilt, taskBottleOri);

  if (glasInHand) {
    for (const std::string& task : {taskGlasOri, taskGlasPosX, taskGlasPosZ}) {
      a1.addActivation(t_start, true, 0.5, task);
      a1.addActivation(t_end + afterTime, false, 0.5, task);
    }
  }
  return a1;
}

}   // namespace affgpu
