#include "ActionScene.h"

#include <algorithm>
#include <cctype>
#include <cmath>
#include <sstream>
#include <stdexcept>

namespace affgpu
{

static bool iequals(const std::string& a, const std::string& b)
{
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}

bool Affordance::typeFromString(const std::string& s, Type& out)
{
  static const std::pair<const char*, Type> tbl[] = {
    {"Affordance", Type::Affordance}, {"Graspable", Type::Graspable}, {"PowerGraspable", Type::PowerGraspable},
    {"PincerGraspable", Type::PincerGraspable}, {"PalmGraspable", Type::PalmGraspable}, {"BallGraspable", Type::BallGraspable},
    {"CircularGraspable", Type::CircularGraspable}, {"TwistGraspable", Type::TwistGraspable}, {"Twistable", Type::Twistable},
    {"PushSwitchable", Type::PushSwitchable}, {"Supportable", Type::Supportable}, {"Stackable", Type::Stackable},
    {"Containable", Type::Containable}, {"Pourable", Type::Pourable}, {"PointPushable", Type::PointPushable},
    {"PointPokable", Type::PointPokable}, {"Hingeable", Type::Hingeable}, {"Dispensible", Type::Dispensible},
    {"Wettable", Type::Wettable}, {"Openable", Type::Openable}};
  for (const auto& e : tbl) if (s == e.first) { out = e.second; return true; }
  return false;
}

// Class hierarchy of reference src/Affordance.h:99-316.
bool Affordance::isA(Type base) const
{
  if (base == Type::Affordance || base == classType) return true;
  if (base == Type::Graspable)
    return classType == Type::PowerGraspable || classType == Type::PincerGraspable || classType == Type::PalmGraspable ||
           classType == Type::BallGraspable || classType == Type::CircularGraspable || classType == Type::TwistGraspable ||
           classType == Type::Twistable;
  if (base == Type::TwistGraspable) return classType == Type::Twistable;
  return false;
}

bool AffordanceEntity::isCollideable(const Graph& graph) const
{
  const int b = graph.getBodyByName(bdyName);
  return b >= 0 && graph.numDistanceShapes(b) > 0;
}

std::vector<double> Manipulator::fingerAnglesFromFingerTipDistance(double d) const
{
  return std::vector<double>(3, 1.0 - d / 0.175);   // reference src/Manipulator.cpp:296-299
}

// reference src/Manipulator.cpp:247-290: a hand is empty unless one of its
// descendant bodies is neither a pure frame nor one of its own grasp frames.
bool Manipulator::isEmpty(const Graph& graph) const
{
  const int hand = graph.getBodyByName(name);
  if (hand < 0) return true;
  for (size_t i = 0; i < graph.bodies.size(); ++i) {
    if ((int)i != hand && !graph.isChild((int)i, hand)) continue;
    const Body& b = graph.bodies[i];
    // The flattener drops FRAME shapes, so "only a frame shape" reads as "no primitive shape".
    if (b.nShapes == 0 && !b.rigid_body_joints) continue;
    if ((int)i == hand) continue;
    bool own = false;
    for (const auto& c : capabilities) if (c.isGrasp() && b.name == c.frame) { own = true; break; }
    if (!own) return false;
  }
  return true;
}

const ModelState* ActionScene::getModelState(const std::string& name, int timeStamp) const
{
  for (const ModelState& m : modelStates) if (m.name == name && m.timeStamp == timeStamp) return &m;
  return nullptr;
}

void ActionScene::parse(const std::vector<std::string>& lines)
{
  AffordanceEntity* ent = nullptr;
  Manipulator* man = nullptr;
  for (const std::string& line : lines) {
    std::istringstream is(line);
    std::string tok;
    is >> tok;
    if (tok == "entities") { size_t n; is >> n; entities.reserve(n); }
    else if (tok == "manipulators") { size_t n; is >> n; manipulators.reserve(n); }
    else if (tok == "entity") {
      AffordanceEntity e; size_t n;
      is >> e.name >> e.bdyName >> e.type >> n;
      entities.push_back(e); ent = &entities.back();
    }
    else if (tok == "aff") {
      if (!ent) throw std::runtime_error("aff line outside entity");
      std::string cls; Affordance a; size_t nkv;
      is >> cls >> a.frame >> nkv;
      if (!Affordance::typeFromString(cls, a.classType)) continue;   // unknown class: ignored like an unknown XML tag
      for (size_t k = 0; k < nkv; ++k) {
        std::string kv; is >> kv;
        const size_t eq = kv.find('=');
        if (eq != std::string::npos) a.attributes[kv.substr(0, eq)] = kv.substr(eq + 1);
      }
      auto num = [&](const char* key, double& v) { auto it = a.attributes.find(key); if (it != a.attributes.end()) v = std::stod(it->second); };
      num("radius", a.radius); num("extentsX", a.extentsX); num("extentsY", a.extentsY);
      auto nd = a.attributes.find("normalDir");
      if (nd != a.attributes.end()) a.normalDir = nd->second == "X" ? 0 : (nd->second == "Y" ? 1 : 2);
      switch (a.classType) {   // reference src/Affordance.cpp:353,423,452,481-482,497
        case Affordance::Type::Stackable: a.requiredAffordances = {Affordance::Type::Supportable}; break;
        case Affordance::Type::Pourable: a.requiredAffordances = {Affordance::Type::Containable}; break;
        case Affordance::Type::PointPokable: a.requiredAffordances = {Affordance::Type::PointPushable}; break;
        case Affordance::Type::Dispensible: a.requiredAffordances = {Affordance::Type::Containable, Affordance::Type::Supportable}; break;
        case Affordance::Type::Wettable: a.requiredAffordances = {Affordance::Type::Containable}; break;
        default: break;
      }
      ent->affordances.push_back(a);
    }
    else if (tok == "manipulator") {
      Manipulator m; size_t nf, nc;
      is >> m.name >> m.type >> nf >> nc;
      manipulators.push_back(m); man = &manipulators.back();
    }
    else if (tok == "modelstate") { ModelState m; size_t n; is >> m.name >> m.timeStamp >> n; modelStates.push_back(m); }
    else if (tok == "msjoint") { std::string j; double v; is >> j >> v; if (!modelStates.empty()) modelStates.back().joints.emplace_back(j, v); }
    else if (tok == "finger") { std::string j; is >> j; if (man) man->fingerJoints.push_back(j); }
    else if (tok == "cap") {
      if (!man) throw std::runtime_error("cap line outside manipulator");
      Capability c;
      is >> c.className >> c.frame;
      using AT = Affordance::Type; using CT = Capability::Type;
      // reference src/Capability.cpp:151-255
      if (c.className == "PowergraspCapability") { c.classType = CT::PowergraspCapability; c.affordanceTypes = {AT::PowerGraspable, AT::Hingeable}; }
      else if (c.className == "PincergraspCapability") { c.classType = CT::PincergraspCapability; c.affordanceTypes = {AT::PincerGraspable, AT::BallGraspable}; }
      else if (c.className == "TwistgraspCapability") { c.classType = CT::TwistgraspCapability; c.affordanceTypes = {AT::TwistGraspable}; }
      else if (c.className == "PalmgraspCapability") { c.classType = CT::PalmgraspCapability; c.affordanceTypes = {AT::PalmGraspable, AT::Hingeable}; }
      else if (c.className == "CirculargraspCapability") {
        c.classType = CT::CirculargraspCapability; c.affordanceTypes = {AT::CircularGraspable};
        const size_t comma = c.frame.find(',');
        if (comma != std::string::npos) { c.frame2 = c.frame.substr(comma + 1); c.frame = c.frame.substr(0, comma); }
      }
      else if (c.className == "FingerpushCapability") { c.classType = CT::FingerpushCapability; c.affordanceTypes = {AT::PushSwitchable}; }
      else if (c.className == "GazeCapability") { c.classType = CT::GazeCapability; }
      man->capabilities.push_back(c);
    }
  }
}

const AffordanceEntity* ActionScene::getAffordanceEntity(const std::string& name) const
{
  for (const auto& e : entities) if (iequals(e.name, name)) return &e;
  return nullptr;
}

std::vector<const AffordanceEntity*> ActionScene::getAffordanceEntities(const std::string& name) const
{
  std::vector<const AffordanceEntity*> found;
  if (name.empty()) return found;
  for (const auto& e : entities) if (iequals(e.name, name)) found.push_back(&e);
  return found;
}

const Manipulator* ActionScene::getManipulator(const std::string& name) const
{
  for (const auto& m : manipulators) if (iequals(m.name, name)) return &m;
  return nullptr;
}

const Manipulator* ActionScene::getManipulator(const Capability* capability) const
{
  for (const auto& m : manipulators)
    for (const auto& c : m.capabilities) if (&c == capability) return &m;
  return nullptr;
}

std::vector<const Manipulator*> ActionScene::getFreeManipulators(const Graph& graph) const
{
  std::vector<const Manipulator*> r;
  for (const auto& m : manipulators) if (m.isEmpty(graph)) r.push_back(&m);
  return r;
}

std::vector<const Manipulator*> ActionScene::getOccupiedManipulators(const Graph& graph) const
{
  std::vector<const Manipulator*> r;
  for (const auto& m : manipulators) if (!m.isEmpty(graph)) r.push_back(&m);
  return r;
}

const Manipulator* ActionScene::getGraspingHand(const Graph& graph, const AffordanceEntity* entity) const
{
  if (!entity) return nullptr;
  const int obj = graph.getBodyByName(entity->bdyName);
  if (obj < 0) return nullptr;
  for (const auto& hand : manipulators)
    for (const auto& c : hand.capabilities) {
      if (!c.isGrasp()) continue;
      const int fr = graph.getBodyByName(c.frame);
      if (fr >= 0 && graph.isChild(obj, fr)) return &hand;
    }
  return nullptr;
}

const AffordanceEntity* ActionScene::getParentAffordanceEntity(const Graph& graph, const AffordanceEntity* child) const
{
  const int b = graph.getBodyByName(child->bdyName);
  for (int p = (b >= 0 ? graph.bodies[b].parent : -1); p >= 0; p = graph.bodies[p].parent)
    for (const auto& e : entities) if (e.bdyName == graph.bodies[p].name) return &e;
  return nullptr;
}

std::vector<AffCapPair> match(Affordance::Type T, const AffordanceEntity* object, const Manipulator* hand)
{
  std::vector<AffCapPair> res;
  for (const Affordance& a : object->affordances) {
    if (!a.isA(T)) continue;
    for (const Capability& c : hand->capabilities)
      for (Affordance::Type ct : c.affordanceTypes)
        if (ct == a.classType) res.emplace_back(&a, &c);
  }
  return res;
}

std::vector<AffAffPair> match(Affordance::Type T1, Affordance::Type T2, const AffordanceEntity* object1, const AffordanceEntity* object2)
{
  std::vector<AffAffPair> res;
  for (const Affordance& a2 : object2->affordances) {
    if (!a2.isA(T2)) continue;
    for (const Affordance& a1 : object1->affordances) {
      if (!a1.isA(T1)) continue;
      for (Affordance::Type req : a2.requiredAffordances)
        if (req == a1.classType) res.emplace_back(&a1, &a2);
    }
  }
  return res;
}

static double frameCost(const Graph& graph, const std::string& f0, const std::string& f1, double wLin, double wAng)
{
  const int b0 = graph.getBodyByName(f0), b1 = graph.getBodyByName(f1);
  if (b0 < 0 || b1 < 0) return 1.0e30;
  const HTr& A = graph.bodies[b0].A_BI; const HTr& B = graph.bodies[b1].A_BI;
  double d2 = 0.0, tr = 0.0;
  for (int k = 0; k < 3; ++k) { const double d = A.org[k] - B.org[k]; d2 += d * d; }
  for (int i = 0; i < 3; ++i) for (int k = 0; k < 3; ++k) tr += A.rot[i][k] * B.rot[i][k];   // trace(A B^T)
  const double c = std::max(-1.0, std::min(1.0, 0.5 * (tr - 1.0)));
  return wLin * std::sqrt(d2) + wAng * std::acos(c);
}

void sort(const Graph& graph, std::vector<AffCapPair>& pairs, double wLin, double wAng)
{
  std::stable_sort(pairs.begin(), pairs.end(), [&](const AffCapPair& a, const AffCapPair& b) {
    return frameCost(graph, std::get<0>(a)->frame, std::get<1>(a)->frame, wLin, wAng) <
           frameCost(graph, std::get<0>(b)->frame, std::get<1>(b)->frame, wLin, wAng);
  });
}

void sort(const Graph& graph, std::vector<AffAffPair>& pairs, double wLin, double wAng)
{
  std::stable_sort(pairs.begin(), pairs.end(), [&](const AffAffPair& a, const AffAffPair& b) {
    return frameCost(graph, std::get<0>(a)->frame, std::get<1>(a)->frame, wLin, wAng) <
           frameCost(graph, std::get<0>(b)->frame, std::get<1>(b)->frame, wLin, wAng);
  });
}

}   // namespace affgpu
