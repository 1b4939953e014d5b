#include "ActionGet.h"

#include <algorithm>
#include <cmath>

namespace affgpu
{

// reference src/ActionGet.cpp:54-60
static const double kFingersOpen = 0.01;
static const double kFingersHalfClosed = 0.7;
static const double kFingersClosed = 0.7;
static const double kFingerMoveTime = 2.0;
static const double kDefaultLiftHeight = 0.12;
static const double kDefaultPreGraspDist = 0.2;
static const double kLiftSafetyDistance = 0.05;
static const double kAfterTime = 0.5;

static double deg2rad(double d) { return d * (M_PI / 180.0); }

static bool takeOption(std::vector<std::string>& params, const std::string& key, std::string& value)
{
  auto it = std::find(params.begin(), params.end(), key);
  if (it == params.end() || it + 1 == params.end()) return false;
  value = *(it + 1);
  params.erase(it, it + 2);
  return true;
}

const char* ActionGet::graspTypeToString(GraspType g)
{
  switch (g) {
    case GraspType::PowerGrasp: return "PowerGrasp";
    case GraspType::BallGrasp: return "BallGrasp";
    case GraspType::TopGrasp: return "TopGrasp";
    case GraspType::RimGrasp: return "RimGrasp";
    case GraspType::OrientedGrasp: return "OrientedGrasp";
    default: return "Other";
  }
}

ActionGet::ActionGet(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
  : liftHeight(kDefaultLiftHeight), preGraspDist(kDefaultPreGraspDist), graspType(GraspType::Other),
    shoulderBase(0.0), handOver(false), isObjCollidable(false)
{
  std::string v;
  if (takeOption(params, "from", v)) whereFrom = v;
  if (takeOption(params, "liftHeight", v)) liftHeight = std::stod(v);
  if (takeOption(params, "duration", v)) defaultDuration = std::stod(v);
  if (params.empty()) throw ActionException("ERROR: REASON: get needs an object", ActionException::ParamNotFound);
  init(domain, graph, params[0], params.size() > 1 ? params[1] : std::string(), whereFrom);
}

void ActionGet::init(const ActionScene& domain, const Graph& graph, const std::string& objectToGet,
                     const std::string& manipulator, const std::string& from)
{
  const std::string err = "ERROR: ";
  const AffordanceEntity* entityToGet = nullptr;
  const Manipulator* specifiedHand = domain.getManipulator(manipulator);
  if (!manipulator.empty() && !specifiedHand)
    throw ActionException(err + " REASON: Can't find a hand with the name " + manipulator +
                          " SUGGESTION: Use a hand name that is defined in the environment", ActionException::ParamNotFound);

  if (from.empty()) {
    const auto ntts = domain.getAffordanceEntities(objectToGet);
    if (ntts.empty())
      throw ActionException(err + " REASON: The " + objectToGet +
                            " is unknown. SUGGESTION: Use an object name that is defined in the environment", ActionException::ParamNotFound);
    // The first entity of that name, unless it is already held as requested.
    const Manipulator* grasping = domain.getGraspingHand(graph, ntts[0]);
    if (grasping && (grasping == specifiedHand || !specifiedHand))
      throw ActionException("SUCCESS DEVELOPER: Object is already held in the hand", ActionException::NoError);
    entityToGet = ntts[0];
  } else {
    std::string fromBody;
    if (const AffordanceEntity* fromNtt = domain.getAffordanceEntity(from)) fromBody = fromNtt->bdyName;
    else if (const Manipulator* fromHand = domain.getManipulator(from)) fromBody = fromHand->name;
    else
      throw ActionException(err + " REASON: The " + from + " is neither a hand nor a part of the environment" +
                            " SUGGESTION: Use an object name or a hand that is defined in the environment", ActionException::ParamNotFound);
    const int bdyFrom = graph.getBodyByName(fromBody);
    for (const AffordanceEntity* ntt : domain.getAffordanceEntities(objectToGet))
      if (graph.isChild(graph.getBodyByName(ntt->bdyName), bdyFrom)) entityToGet = ntt;
    if (!entityToGet)
      throw ActionException(err + " REASON: The " + from + " is empty", ActionException::ParamNotFound);
  }

  isObjCollidable = entityToGet->isCollideable(graph);

  bool graspable = false;
  for (const Affordance& a : entityToGet->affordances) graspable = graspable || a.isA(Affordance::Type::Graspable);
  if (!graspable) throw ActionException(err + " REASON: " + objectToGet + " cannot be grasped", ActionException::ParamNotFound);

  const int objBdy = graph.getBodyByNameNoCase(entityToGet->bdyName);
  if (objBdy < 0) throw ActionException(err + "object body " + entityToGet->bdyName + " is not part of the graph", ActionException::ParamNotFound);
  objectName = graph.bodies[objBdy].name;

  // Lift height: limited by whatever is above the object (src/ActionGet.cpp:241-266).
  double mn[3], mx[3], castFrom[3], dMin = 0.0;
  if (graph.computeBodyAABB(objBdy, mn, mx)) { castFrom[0] = 0.5 * (mn[0] + mx[0]); castFrom[1] = 0.5 * (mn[1] + mx[1]); castFrom[2] = mx[2] + 1.0e-8; }
  else for (int k = 0; k < 3; ++k) castFrom[k] = graph.bodies[objBdy].A_BI.org[k];
  const double up[3] = {0.0, 0.0, 1.0};
  const int above = graph.closestRigidBodyInDirection(castFrom, up, &dMin);
  if (above >= 0 && !graph.isChild(above, objBdy))
    liftHeight = std::max(0.0, std::min(liftHeight, std::min(liftHeight, dMin - kLiftSafetyDistance)));
  liftHeight += graph.bodies[objBdy].A_BI.org[2];   // absolute coordinates

  std::vector<const Manipulator*> freeManipulators;
  if (manipulator.empty()) freeManipulators = domain.getFreeManipulators(graph);
  else {
    if (!specifiedHand->isEmpty(graph))
      throw ActionException(err + " REASON: The hand " + manipulator + " is already holding something" +
                            " SUGGESTION: Free your hand before performing this command, or use another hand", ActionException::ParamNotFound);
    freeManipulators.push_back(specifiedHand);
  }
  if (freeManipulators.empty())
    throw ActionException(err + " REASON: All hands are already full SUGGESTION: Free one or all your hands before performing this command",
                          ActionException::KinematicallyImpossible);

  for (const Manipulator* hand : freeManipulators) {
    const auto m = match(Affordance::Type::Graspable, entityToGet, hand);
    affordanceMap.insert(affordanceMap.end(), m.begin(), m.end());
  }
  if (affordanceMap.empty())
    throw ActionException(err + " REASON: I don't have the capability of grasping the " + objectToGet, ActionException::KinematicallyImpossible);

  sort(graph, affordanceMap, 1.0, 0.0);

  if (const Manipulator* graspingHand = domain.getGraspingHand(graph, entityToGet)) {
    handOver = true;
    handOverHand = graspingHand->name;
    liftHeight = graph.bodies[objBdy].A_BI.org[2];
  }
  if (!initialize(domain, graph, 0)) throw ActionException(err + "could not initialize solution 0", ActionException::UnknownError);
}

bool ActionGet::initialize(const ActionScene& domain, const Graph& graph, size_t rank)
{
  if (rank >= getNumSolutions()) return false;
  const Capability* cap = std::get<1>(affordanceMap[rank]);
  const Affordance* aff = std::get<0>(affordanceMap[rank]);
  const Manipulator* hand = domain.getManipulator(cap);
  if (!hand) return false;
  capabilityFrame = cap->frame;
  affordanceFrame = aff->frame;
  handOpen.assign(hand->getNumFingers(), kFingersOpen);
  handClosed.assign(hand->getNumFingers(), kFingersClosed);

  using AT = Affordance::Type;
  if (aff->classType == AT::PowerGraspable) graspType = GraspType::PowerGrasp;
  else if (aff->classType == AT::BallGraspable) {
    handOpen.assign(hand->getNumFingers(), kFingersHalfClosed);
    handClosed = hand->fingerAnglesFromFingerTipDistance(2.0 * aff->radius);
    graspType = GraspType::BallGrasp;
  }
  else if (aff->isA(AT::TwistGraspable)) {
    handClosed = hand->fingerAnglesFromFingerTipDistance(2.0 * aff->radius);
    graspType = GraspType::TopGrasp;
  }
  else if (aff->classType == AT::CircularGraspable) graspType = GraspType::RimGrasp;
  else graspType = GraspType::OrientedGrasp;

  // Shoulder side: walk up from the hand until the parent branches (src/ActionGet.cpp:441-458).
  shoulderBase = 0.0;
  {
    int b = graph.getBodyByName(hand->name);
    if (b < 0) return false;
    b = graph.bodies[b].parent;
    while (b >= 0 && graph.bodies[b].parent != 1) {
      const int p = graph.bodies[b].parent;
      if (p < 0 || graph.children(p).size() > 1) { shoulderBase = graph.bodies[b].A_BI.org[1]; break; }
      b = p;
    }
  }

  usedManipulators.assign(1, hand->name);
  for (const Manipulator* m : domain.getOccupiedManipulators(graph)) if (m->name != hand->name) usedManipulators.push_back(m->name);
  fingerJoints = hand->fingerJoints;

  taskObjHandPos = objectName + "-" + capabilityFrame + "-XYZ";
  taskObjPosX = objectName + "-X";
  taskObjPosY = objectName + "-Y";
  taskObjPosZ = objectName + "-Z";
  taskObjOri = objectName + "-POLAR";
  taskHandObjOri = capabilityFrame + "-" + objectName + "-" + graspTypeToString(graspType);
  taskFingers = capabilityFrame + "_fingers";
  taskHandOverHand = handOverHand.empty() ? std::string() : handOverHand + "_XYZ";
  return true;
}

std::vector<TaskSpec> ActionGet::createTasks() const
{
  if (graspType == GraspType::RimGrasp || graspType == GraspType::OrientedGrasp)
    throw ActionException(std::string("ERROR: grasp type ") + graspTypeToString(graspType) +
                          " needs Composite/Radial/ABC tasks, which the GPU predictor does not cover yet", ActionException::ParamInvalid);
  std::vector<TaskSpec> tasks;
  auto mk = [](const std::string& name, const std::string& cv, const std::string& eff, const std::string& ref = "",
               const std::string& axis = "") {
    TaskSpec t; t.name = name; t.controlVariable = cv; t.effector = eff; t.refBdy = ref; t.axisDirection = axis; return t;
  };
  if (graspType == GraspType::BallGrasp || graspType == GraspType::TopGrasp)
    tasks.push_back(mk(taskObjHandPos, "XYZ", capabilityFrame, affordanceFrame));   // hand relative to the object
  else
    tasks.push_back(mk(taskObjHandPos, "XYZ", affordanceFrame, capabilityFrame));   // object relative to the hand
  tasks.push_back(mk(taskObjPosX, "X", objectName));
  tasks.push_back(mk(taskObjPosY, "Y", objectName));
  tasks.push_back(mk(taskObjPosZ, "Z", objectName));
  tasks.push_back(mk(taskObjOri, "POLAR", affordanceFrame));
  if (graspType == GraspType::PowerGrasp) tasks.push_back(mk(taskHandObjOri, "POLAR", capabilityFrame, affordanceFrame));
  else if (graspType == GraspType::TopGrasp) tasks.push_back(mk(taskHandObjOri, "POLAR", capabilityFrame, affordanceFrame, "X"));
  else tasks.push_back(mk(taskHandObjOri, "Inclination", capabilityFrame, "", "X"));
  if (!taskHandOverHand.empty()) tasks.push_back(mk(taskHandOverHand, "XYZ", handOverHand));
  TaskSpec fingers; fingers.name = taskFingers; fingers.controlVariable = "Joints"; fingers.jnts = fingerJoints;
  tasks.push_back(fingers);
  return tasks;
}

ActivationSet ActionGet::createTrajectory(double t_start, double t_end) const
{
  const double t_grasp = t_start + 0.66 * (t_end - t_start);
  const double t_pregrasp = t_start + 0.5 * (t_grasp - t_start);
  ActivationSet a1;
  if (graspType == GraspType::BallGrasp) a1 = trajectoryBallGrasp(t_start, t_grasp, t_end);
  else if (graspType == GraspType::TopGrasp) a1 = trajectoryTopGrasp(t_start, t_grasp, t_end);
  else a1 = trajectoryOriented(t_start, t_grasp, t_end);

  if (!handOpen.empty() && !handClosed.empty()) {
    a1.addActivation(t_start, true, 0.5, taskFingers);
    a1.addVectorConstraint(t_grasp - 0.5 * kFingerMoveTime, handOpen, taskFingers);
    a1.addVectorConstraint(t_grasp + 0.5 * kFingerMoveTime, handClosed, taskFingers);
  }
  if (isObjCollidable) {
    a1.addCollisionModelConstraint(t_pregrasp, objectName, false);   // no object collisions while approaching
    a1.addCollisionModelConstraint(t_end, objectName, true);         // back on once lifted
  }
  return a1;
}

// Lift block shared by the grasp variants: keep the object upright and move it
// up / towards the body after t_on.
static void addLift(ActivationSet& a1, double t_on, double t_off, const std::string& ori, const std::string& px,
                    const std::string& py, const std::string& pz, bool withOri, bool withX)
{
  if (withOri) { a1.addActivation(t_on, true, 0.5, ori); a1.addActivation(t_off, false, 0.5, ori); }
  a1.addActivation(t_on, true, 0.5, pz); a1.addActivation(t_off, false, 0.5, pz);
  if (withX) { a1.addActivation(t_on, true, 0.5, px); a1.addActivation(t_off, false, 0.5, px); }
  (void)py;
}

ActivationSet ActionGet::trajectoryOriented(double t_start, double t_grasp, double t_end) const
{
  ActivationSet a1;
  const double t_pregrasp = t_start + 0.5 * (t_grasp - t_start);
  a1.addActivation(t_start, true, 0.5, taskObjHandPos);
  a1.addActivation(t_grasp, false, 0.5, taskObjHandPos);
  a1.add(t_pregrasp, preGraspDist, 0.0, 0.0, 1, taskObjHandPos + " 0");   // some distance in front
  a1.add(t_pregrasp, 0.0, 0.0, 0.0, 7, taskObjHandPos + " 1");            // centred
  a1.add(t_pregrasp, 0.0, 0.0, 0.0, 7, taskObjHandPos + " 2");
  a1.addPositionConstraint(t_grasp, 0.0, 0.0, 0.0, taskObjHandPos);

  a1.addActivation(t_start, true, 0.5, taskHandObjOri);
  a1.addActivation(t_grasp, false, 0.5, taskHandObjOri);
  a1.addPolarConstraint(t_grasp, 0.0, 0.0, taskHandObjOri);              // PowerGrasp

  const double t_on = handOver ? t_start : t_grasp;
  addLift(a1, t_on, t_end + kAfterTime, taskObjOri, taskObjPosX, taskObjPosY, taskObjPosZ, true, true);
  a1.add(t_end, liftHeight, 0.0, 0.0, 7, taskObjPosZ + " 0");
  a1.addConnectBodyConstraint(t_grasp, objectName, capabilityFrame);
  a1.addActivation(t_grasp, true, 0.5, taskObjPosY);
  a1.addActivation(t_end + kAfterTime, false, 0.5, taskObjPosY);
  if (shoulderBase != 0.0) {   // retract towards the shoulder side
    a1.add(t_end, -0.35, 0.0, 0.0, 7, taskObjPosX + " 0");
    a1.add(t_end, shoulderBase < 0.0 ? -0.3 : 0.3, 0.0, 0.0, 7, taskObjPosY + " 0");
  }
  if (!taskHandOverHand.empty()) {
    a1.addActivation(t_grasp, true, 0.5, taskHandOverHand);
    a1.addActivation(t_end, false, 0.5, taskHandOverHand);
  }
  return a1;
}

ActivationSet ActionGet::trajectoryBallGrasp(double t_start, double t_grasp, double t_end) const
{
  ActivationSet a1;
  const double t_pregrasp = t_start + 0.5 * (t_grasp - t_start);
  a1.addActivation(t_start, true, 0.5, taskObjHandPos);
  a1.addActivation(t_grasp, false, 0.5, taskObjHandPos);
  a1.addPositionConstraint(t_pregrasp, 0.0, 0.0, preGraspDist, taskObjHandPos);
  a1.addPositionConstraint(t_grasp, 0.0, 0.0, 0.0, taskObjHandPos);

  a1.addActivation(t_start, true, 0.5, taskHandObjOri);
  a1.addActivation(t_end + kAfterTime, false, 0.5, taskHandObjOri);
  a1.add(t_grasp, deg2rad(160.0), 0.0, 0.0, 7, taskHandObjOri + " 0");
  a1.add(t_end, deg2rad(160.0), 0.0, 0.0, 7, taskHandObjOri + " 0");

  addLift(a1, t_grasp, t_end + kAfterTime, taskObjOri, taskObjPosX, taskObjPosY, taskObjPosZ, false, !handOver);
  a1.add(t_end, liftHeight, 0.0, 0.0, 7, taskObjPosZ + " 0");
  a1.addConnectBodyConstraint(t_grasp, objectName, capabilityFrame);
  a1.addActivation(t_grasp, true, 0.5, taskObjPosY);
  a1.addActivation(t_end + kAfterTime, false, 0.5, taskObjPosY);
  return a1;
}

ActivationSet ActionGet::trajectoryTopGrasp(double t_start, double t_grasp, double t_end) const
{
  ActivationSet a1;
  const double t_pregrasp = t_start + 0.5 * (t_grasp - t_start);
  a1.addActivation(t_start, true, 0.5, taskObjHandPos);
  a1.addActivation(t_grasp, false, 0.5, taskObjHandPos);
  a1.addPositionConstraint(t_pregrasp, 0.0, 0.0, 0.5 * preGraspDist, taskObjHandPos);
  a1.addPositionConstraint(t_grasp, 0.0, 0.0, 0.0, taskObjHandPos);

  a1.addActivation(t_start, true, 0.5, taskHandObjOri);
  a1.addActivation(t_grasp, false, 0.5, taskHandObjOri);
  a1.addPolarConstraint(t_grasp, M_PI, 0.0, taskHandObjOri);
  a1.addPolarConstraint(t_end, M_PI, 0.0, taskHandObjOri);

  const double t_on = handOver ? t_start : t_grasp;
  addLift(a1, t_on, t_end + kAfterTime, taskObjOri, taskObjPosX, taskObjPosY, taskObjPosZ, true, true);
  if (!handOver) a1.add(t_end, liftHeight, 0.0, 0.0, 7, taskObjPosZ + " 0");
  a1.addConnectBodyConstraint(t_grasp, objectName, capabilityFrame);
  a1.addActivation(t_grasp, true, 0.5, taskObjPosY);
  a1.addActivation(t_end + kAfterTime, false, 0.5, taskObjPosY);
  a1.add(t_end, -0.35, 0.0, 0.0, 7, taskObjPosX + " 0");
  a1.add(t_end, 0.0, 0.0, 0.0, 7, taskObjPosY + " 0");
  if (!taskHandOverHand.empty()) {
    a1.addActivation(t_grasp, true, 0.5, taskHandOverHand);
    a1.addActivation(t_end, false, 0.5, taskHandOverHand);
  }
  return a1;
}

}   // namespace affgpu
