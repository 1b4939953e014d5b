// Symbolic scene layer that feeds the candidate lists of the predict path:
// affordances, capabilities, manipulators and their matching
// (reference src/ActionScene.h:116-226, src/Affordance.h, src/Capability.cpp:151-255,
//  src/Manipulator.cpp, src/AffordanceEntity.cpp:271-443).
#ifndef AFF_HOST_ACTIONSCENE_H
#define AFF_HOST_ACTIONSCENE_H

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "Graph.h"

namespace affgpu
{

struct Affordance
{
  // Same names as the reference's Affordance::Type enumerators.
  enum class Type { Affordance, Graspable, PowerGraspable, PincerGraspable, PalmGraspable, BallGraspable,
                    CircularGraspable, TwistGraspable, Twistable, PushSwitchable, Supportable, Stackable,
                    Containable, Pourable, PointPushable, PointPokable, Hingeable, Dispensible, Wettable, Openable };
  static bool typeFromString(const std::string& s, Type& out);
  // dynamic_cast<T*>(aff) in the reference == isA(T)
  bool isA(Type base) const;

  Type classType = Type::Affordance;
  std::string frame;
  std::vector<Type> requiredAffordances;
  double radius = 0.0;                  // BallGraspable / TwistGraspable
  double extentsX = 0.0, extentsY = 0.0; // Supportable
  int normalDir = 2;                    // Stackable
  std::map<std::string, std::string> attributes;
};

struct AffordanceEntity
{
  std::string name, bdyName, type;
  std::vector<Affordance> affordances;
  bool isCollideable(const Graph& graph) const;
};

struct Capability
{
  enum class Type { Capability, GraspCapability, PowergraspCapability, PincergraspCapability, TwistgraspCapability,
                    CirculargraspCapability, PalmgraspCapability, FingerpushCapability, GazeCapability };
  Type classType = Type::Capability;
  std::string className, frame, frame2;
  std::vector<Affordance::Type> affordanceTypes;
  bool isGrasp() const { return classType != Type::Capability && classType != Type::FingerpushCapability && classType != Type::GazeCapability; }
};

struct Manipulator
{
  std::string name, type;
  std::vector<std::string> fingerJoints;
  std::vector<Capability> capabilities;

  size_t getNumFingers() const { return fingerJoints.size(); }
  std::vector<double> fingerAnglesFromFingerTipDistance(double fingerTipDistanceInMeters) const;
  bool isEmpty(const Graph& graph) const;
};

// One <model_state model=".." time_stamp=".."> of the graph file: joint name -> value (rad / m).
struct ModelState
{
  std::string name;
  int timeStamp = 0;
  std::vector<std::pair<std::string, double>> joints;
};

using AffCapPair = std::tuple<const Affordance*, const Capability*>;
using AffAffPair = std::tuple<const Affordance*, const Affordance*>;

class ActionScene
{
public:
  // Parses the entity / manipulator tail of an .affscene file.
  void parse(const std::vector<std::string>& lines);

  const AffordanceEntity* getAffordanceEntity(const std::string& name) const;
  std::vector<const AffordanceEntity*> getAffordanceEntities(const std::string& name) const;
  const Manipulator* getManipulator(const std::string& name) const;
  const Manipulator* getManipulator(const Capability* capability) const;
  std::vector<const Manipulator*> getFreeManipulators(const Graph& graph) const;
  std::vector<const Manipulator*> getOccupiedManipulators(const Graph& graph) const;
  const Manipulator* getGraspingHand(const Graph& graph, const AffordanceEntity* entity) const;
  const AffordanceEntity* getParentAffordanceEntity(const Graph& graph, const AffordanceEntity* child) const;

  // RcsGraph_getModelStateFromXML: the named state, or null
  const ModelState* getModelState(const std::string& name, int timeStamp) const;

  std::vector<AffordanceEntity> entities;
  std::vector<Manipulator> manipulators;
  std::vector<ModelState> modelStates;
};

// match<T>(object, hand): every (affordance of type T or derived) x (capability
// listing exactly that affordance's class) pair  (reference src/ActionScene.h:116-160).
std::vector<AffCapPair> match(Affordance::Type T, const AffordanceEntity* object, const Manipulator* hand);
// match<T1,T2>(object1, object2)  (reference src/ActionScene.h:181-226).
std::vector<AffAffPair> match(Affordance::Type T1, Affordance::Type T2, const AffordanceEntity* object1, const AffordanceEntity* object2);

// Sort ascending by wLin*|d org| + wAng*diffAngle(rot) between the two frames
// (reference src/AffordanceEntity.cpp:280-443).  Stable, so equal costs keep
// their match order (std::sort in the reference leaves that unspecified).
void sort(const Graph& graph, std::vector<AffCapPair>& pairs, double wLin, double wAng);
void sort(const Graph& graph, std::vector<AffAffPair>& pairs, double wLin, double wAng);

}   // namespace affgpu

#endif
