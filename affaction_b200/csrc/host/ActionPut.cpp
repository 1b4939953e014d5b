#include "ActionPut.h"

#include <algorithm>
#include <cfloat>
#include <cmath>

namespace affgpu
{

static const double kFingersOpen = 0.01;      // reference src/ActionPut.cpp:50-ff (same values as ActionGet)
static const double kFingerMoveTime = 2.0;
static const double kAfterTime = 0.5;

static bool takeOption(std::vector<std::string>& params, const std::string& key, std::string& value)
{
  auto it = std::find(params.begin(), params.end(), key);
  if (it == params.end() || it + 1 == params.end()) return false;
  value = *(it + 1);
  params.erase(it, it + 2);
  return true;
}

ActionPut::ActionPut(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
  : putDown(false), isObjCollidable(false), isPincerGrasped(false), supportRegionX(0.0), supportRegionY(0.0), polarAxisIdx(2)
{
  downProjection[0] = downProjection[1] = downProjection[2] = 0.0;
  std::string v;
  if (takeOption(params, "frame", v)) whereOn = v;
  if (takeOption(params, "duration", v)) defaultDuration = std::stod(v);
  if (params.empty()) throw ActionException("ERROR REASON: put needs an object", ActionException::ParamNotFound);
  init(domain, graph, params[0], params.size() > 1 ? params[1] : std::string());
}

// Manipulator::getGrasp (reference src/Manipulator.cpp:301-345): the object's affordance frame closest to
// the grasp frame.  The reference's angular term reads (org[2], rot[0][0], rot[0][1]) of each frame as a
// vector (it takes the address of org[2]); that is restated as written.
static const Affordance* graspedAffordance(const Graph& graph, const Capability* cap, const AffordanceEntity* entity)
{
  const int gf = graph.getBodyByName(cap->frame);
  const HTr& G = graph.bodies[gf].A_BI;
  double best = DBL_MAX;
  const Affordance* res = nullptr;
  for (const Affordance& a : entity->affordances) {
    const int af = graph.getBodyByName(a.frame);
    if (af < 0) continue;
    const HTr& A = graph.bodies[af].A_BI;
    const double u[3] = {G.org[2], G.rot[0][0], G.rot[0][1]}, v[3] = {A.org[2], A.rot[0][0], A.rot[0][1]};
    const double nu = std::sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]), nv = std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    double dist = 0.0;
    if (nu > 0.0 && nv > 0.0) dist = std::acos(std::max(-1.0, std::min(1.0, (u[0] * v[0] + u[1] * v[1] + u[2] * v[2]) / (nu * nv))));
    double d2 = 0.0;
    for (int k = 0; k < 3; ++k) d2 += (G.org[k] - A.org[k]) * (G.org[k] - A.org[k]);
    dist += std::sqrt(d2);
    if (dist < best) { best = dist; res = &a; }
  }
  return res;
}

void ActionPut::init(const ActionScene& domain, const Graph& graph, const std::string& objectToPut, const std::string& surfaceToPutOn)
{
  const auto objects = domain.getAffordanceEntities(objectToPut);
  if (objects.empty())
    throw ActionException("ERROR REASON: The " + objectToPut + " is unknown. SUGGESTION: Use an object name that is defined in the environment",
                          ActionException::ParamNotFound);
  const Manipulator* graspingHand = nullptr;
  const AffordanceEntity* object = nullptr;
  for (const AffordanceEntity* e : objects) if ((graspingHand = domain.getGraspingHand(graph, e))) { object = e; break; }
  if (!graspingHand)
    throw ActionException("ERROR REASON: The " + objectToPut + " is not held in the hand. SUGGESTION: First get the " + objectToPut +
                          " in the hand before performing this command", ActionException::KinematicallyImpossible);
  bool stackable = false;
  for (const Affordance& a : object->affordances) stackable = stackable || a.isA(Affordance::Type::Stackable);
  if (!stackable)
    throw ActionException("ERROR REASON: I don't know how to place the " + objectToPut + " on a surface. SUGGESTION: Try to hand it over, or try something else",
                          ActionException::KinematicallyImpossible);

  objName = object->bdyName;
  isObjCollidable = object->isCollideable(graph);
  usedManipulators.assign(1, graspingHand->name);
  for (const Manipulator* m : domain.getOccupiedManipulators(graph)) if (m->name != graspingHand->name) usedManipulators.push_back(m->name);

  const int objBody = graph.getBodyByName(objName);
  const Capability* graspCap = getGraspingCapability(graph, graspingHand, objBody);
  if (!graspCap) throw ActionException("ERROR REASON: cannot tell how the " + objectToPut + " is held", ActionException::UnknownError);
  isPincerGrasped = graspCap->classType == Capability::Type::PincergraspCapability;
  graspFrame = graspCap->frame;
  const Affordance* aGrasp = graspedAffordance(graph, graspCap, object);
  if (!aGrasp) throw ActionException("ERROR REASON: no grasp frame on the " + objectToPut, ActionException::UnknownError);
  objGraspFrame = aGrasp->frame;

  const AffordanceEntity* surface = domain.getAffordanceEntity(surfaceToPutOn);
  if (!surface && !surfaceToPutOn.empty())
    throw ActionException("ERROR REASON: The surface " + surfaceToPutOn + " to put the " + objectToPut +
                          " on is unknown. SUGGESTION: Use an object name that is defined in the environment", ActionException::ParamNotFound);
  if (!surface) {
    surface = raycastSurface(domain, graph, objBody);
    if (!surface)
      throw ActionException("ERROR REASON: There is nothing under the " + objectToPut + " where I can put it on. SUGGESTION: Specify another object to put it on",
                            ActionException::KinematicallyImpossible);
    putDown = true;
  }

  affordanceMap = match(Affordance::Type::Supportable, Affordance::Type::Stackable, surface, object);
  if (affordanceMap.empty())
    throw ActionException("ERROR REASON: I can't put the " + objectToPut + " on the " + surface->name +
                          " because it does not support it. SUGGESTION: Specify another object to put it on", ActionException::ParamNotFound);
  if (!whereOn.empty())
    affordanceMap.erase(std::remove_if(affordanceMap.begin(), affordanceMap.end(),
                                       [&](const AffAffPair& p) { return std::get<0>(p)->frame != whereOn; }), affordanceMap.end());
  if (affordanceMap.empty())
    throw ActionException("ERROR REASON: I can't put the " + objectToPut + " on the " + surface->name + "'s frame " + whereOn + ".", ActionException::ParamNotFound);

  // Supportables without extents that already carry a collideable entity are taken (src/ActionPut.cpp:228-287)
  affordanceMap.erase(std::remove_if(affordanceMap.begin(), affordanceMap.end(), [&](const AffAffPair& p) {
    const Affordance* s = std::get<0>(p);
    if (s->extentsX > 0.0 || s->extentsY > 0.0) return false;
    const int frame = graph.getBodyByName(s->frame);
    for (int child : graph.children(frame)) {
      const AffordanceEntity* ntt = domain.getAffordanceEntity(graph.bodies[child].name);
      if (ntt && ntt->isCollideable(graph)) return true;
    }
    return false;
  }), affordanceMap.end());
  if (affordanceMap.empty())
    throw ActionException("ERROR REASON: I can't put the " + objectToPut + " on the " + surface->name +
                          ". There is already something on it. SUGGESTION: Put the object somewhere else, or remove the blocking object. ",
                          ActionException::ParamNotFound);

  sort(graph, affordanceMap, 1.0, 1.0);
  if (!initialize(domain, graph, 0)) throw ActionException("ERROR: could not initialize solution 0", ActionException::UnknownError);
}

bool ActionPut::initialize(const ActionScene& domain, const Graph& graph, size_t rank)
{
  if (rank >= affordanceMap.size()) return false;
  const Affordance* surfaceAff = std::get<0>(affordanceMap[rank]);
  const Affordance* bottomAff = std::get<1>(affordanceMap[rank]);
  surfaceFrameName = surfaceAff->frame;
  supportRegionX = surfaceAff->extentsX;
  supportRegionY = surfaceAff->extentsY;
  polarAxisIdx = bottomAff->normalDir;
  objBottomName = bottomAff->frame;
  if (graph.getBodyByName(surfaceFrameName) < 0) return false;
  downProjection[0] = downProjection[1] = downProjection[2] = 0.0;   // surface origin in its own frame (:338)

  const Manipulator* hand = domain.getManipulator(usedManipulators[0]);
  if (!hand) return false;
  fingerJoints = hand->fingerJoints;
  handOpen.assign(hand->getNumFingers(), kFingersOpen);

  taskObjHandPos = objGraspFrame + "-" + graspFrame + "-XYZ";
  taskHandSurfacePos = graspFrame + "-" + surfaceFrameName + "-XYZ";
  taskObjSurfacePosX = objBottomName + "-" + surfaceFrameName + "-X";
  taskObjSurfacePosY = objBottomName + "-" + surfaceFrameName + "-Y";
  taskObjSurfacePosZ = objBottomName + "-" + surfaceFrameName + "-Z";
  taskObjSurfacePolar = objBottomName + "-" + surfaceFrameName + "-POLAR";
  taskHandPolar = graspFrame + "-POLAR";
  taskHandObjPolar = graspFrame + "-" + objBottomName + "-POLAR";
  taskSurfaceOri = surfaceFrameName + "-POLAR";
  taskFingers = graspFrame + "_fingers";
  explanation = "I'm putting the " + objName + " on the " + surfaceFrameName;
  return true;
}

std::vector<TaskSpec> ActionPut::createTasks() const
{
  std::vector<TaskSpec> tasks;
  auto mk = [](const std::string& name, const std::string& cv, const std::string& eff, const std::string& ref = "",
               const std::string& frame = "", const std::string& axis = "") {
    TaskSpec t; t.name = name; t.controlVariable = cv; t.effector = eff; t.refBdy = ref; t.refFrame = frame; t.axisDirection = axis; return t;
  };
  const bool useTaskRegion = (supportRegionX != 0.0) || (supportRegionY != 0.0);
  tasks.push_back(mk(taskObjHandPos, "XYZ", objGraspFrame, graspFrame));
  tasks.push_back(mk(taskHandSurfacePos, "XYZ", graspFrame, objGraspFrame, surfaceFrameName));   // vertical retract
  TaskSpec px = mk(taskObjSurfacePosX, "X", objBottomName, surfaceFrameName);
  TaskSpec py = mk(taskObjSurfacePosY, "Y", objBottomName, surfaceFrameName);
  if (useTaskRegion) {
    // the reference writes the bounds through std::to_string (6 decimals) into the task XML
    px.hasRegion = py.hasRegion = true;
    px.regionMin = std::stod(std::to_string(-0.5 * supportRegionX)); px.regionMax = std::stod(std::to_string(0.5 * supportRegionX));
    py.regionMin = std::stod(std::to_string(-0.5 * supportRegionY)); py.regionMax = std::stod(std::to_string(0.5 * supportRegionY));
    px.dxScaling = py.dxScaling = 0.0;
  }
  tasks.push_back(px);
  tasks.push_back(py);
  tasks.push_back(mk(taskObjSurfacePosZ, "Z", objBottomName, surfaceFrameName));
  tasks.push_back(mk(taskObjSurfacePolar, "POLAR", objBottomName, surfaceFrameName, "", polarAxisIdx == 0 ? "X" : (polarAxisIdx == 1 ? "Y" : "")));
  tasks.push_back(mk(taskHandPolar, "Inclination", graspFrame, "", "", "X"));
  if (isPincerGrasped) tasks.push_back(mk(taskHandObjPolar, "Inclination", graspFrame, objBottomName, "", "X"));
  else tasks.push_back(mk(taskHandObjPolar, "POLAR", graspFrame, objBottomName));
  tasks.push_back(mk(taskSurfaceOri, "POLAR", surfaceFrameName));
  TaskSpec fingers; fingers.name = taskFingers; fingers.controlVariable = "Joints"; fingers.jnts = fingerJoints;
  tasks.push_back(fingers);
  return tasks;
}

ActivationSet ActionPut::createTrajectory(double t_start, double t_release) const
{
  const double t_put = t_start + 0.66 * (t_release - t_start);
  ActivationSet a1;
  a1.addActivation(t_start, true, 0.5, taskSurfaceOri);
  a1.addActivation(t_release, false, 0.5, taskSurfaceOri);
  for (const std::string* t : {&taskObjSurfacePosX, &taskObjSurfacePosY, &taskObjSurfacePosZ}) {
    a1.addActivation(t_start, true, 0.5, *t);
    a1.addActivation(t_put, false, 0.5, *t);
  }
  a1.add(t_put, downProjection[0], 0.0, 0.0, 7, taskObjSurfacePosX + " 0");
  a1.add(t_put, downProjection[1], 0.0, 0.0, 7, taskObjSurfacePosY + " 0");
  a1.add(t_put, downProjection[2], 0.0, 0.0, 7, taskObjSurfacePosZ + " 0");
  if (!putDown) a1.add(0.75 * (t_start + t_put), 0.1, 0.0, 0.0, 1, taskObjSurfacePosZ + " 0");   // a little curve above other objects
  a1.addConnectBodyConstraint(t_put, objName, surfaceFrameName);
  if (isPincerGrasped) {
    a1.addActivation(t_put, true, 0.5, taskHandSurfacePos);
    a1.addActivation(t_release + kAfterTime, false, 0.5, taskHandSurfacePos);
    a1.add(t_release, 0.15, 0.0, 0.0, 7, taskHandSurfacePos + " 2");
    a1.addActivation(t_put, true, 0.5, taskHandPolar);
    a1.addActivation(t_put + 0.5 * (t_release - t_put), false, 0.5, taskHandPolar);
  } else {
    a1.addActivation(t_put, true, 0.5, taskObjHandPos);
    a1.addActivation(t_release, false, 0.5, taskObjHandPos);
    a1.add(t_release, 0.15, 0.0, 0.0, 7, taskObjHandPos + " 0");
    a1.add(t_release, -0.05, 0.0, 0.0, 7, taskObjHandPos + " 2");
  }
  a1.addActivation(t_start, true, 0.5, taskObjSurfacePolar);
  a1.addActivation(t_put, false, 0.5, taskObjSurfacePolar);
  a1.addPolarConstraint(t_put, 0.0, 0.0, taskObjSurfacePolar);
  if (!isPincerGrasped) {
    a1.addActivation(t_put, true, 0.5, taskHandObjPolar);
    a1.addActivation(t_put + 0.5 * (t_release - t_put), false, 0.5, taskHandObjPolar);
  }
  a1.addActivation(t_put - 0.5 * kFingerMoveTime, true, 0.5, taskFingers);
  a1.addActivation(t_release, false, 0.5, taskFingers);
  if (!handOpen.empty()) a1.addVectorConstraint(t_put + 0.5 * kFingerMoveTime, handOpen, taskFingers);
  if (isObjCollidable) {
    a1.addCollisionModelConstraint(t_put, objBottomName, false);
    a1.addCollisionModelConstraint(t_release, objBottomName, true);
  }
  return a1;
}

}   // namespace affgpu
