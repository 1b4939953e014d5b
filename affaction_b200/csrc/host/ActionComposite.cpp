#include "ActionComposite.h"

#include <array>
#include <cfloat>
#include <cmath>

#include "ActionGet.h"

namespace affgpu
{

ActionComposite::ActionComposite(const ActionComposite& other) : ActionBase(other)
{
  for (const auto& a : other.actions) actions.push_back(a->clone());
}

// Duplicate tasks (same name) are dropped when the batch is assembled (CandidateBatch::append,
// reference src/ActionBase.cpp:95-110), so the lists are simply concatenated here.
std::vector<TaskSpec> ActionComposite::createTasks() const
{
  std::vector<TaskSpec> tasks;
  for (const auto& a : actions) {
    const std::vector<TaskSpec> ti = a->createTasks();
    tasks.insert(tasks.end(), ti.begin(), ti.end());
  }
  return tasks;
}

ActivationSet ActionComposite::createTrajectory(double t_start, double t_end) const
{
  ActivationSet cset;
  for (const auto& a : actions) cset.add(a->createTrajectory(t_start, t_end));
  return cset;
}

double ActionComposite::getDurationHint() const
{
  double duration = 0.0;
  for (const auto& a : actions) duration = std::max(duration, a->getDurationHint());
  return duration;
}

std::vector<std::string> ActionComposite::getManipulators() const
{
  std::vector<std::string> mVec;
  for (const auto& a : actions) {
    const std::vector<std::string> mi = a->getManipulators();
    mVec.insert(mVec.end(), mi.begin(), mi.end());
  }
  return mVec;
}

size_t ActionComposite::getNumSolutions() const
{
  size_t n = actions.empty() ? 0 : 1;
  for (const auto& a : actions) n *= a->getNumSolutions();
  return n;
}

bool ActionComposite::initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank)
{
  if (solutionRank >= getNumSolutions()) return false;
  for (size_t i = actions.size(); i-- > 0;) {
    const size_t n = actions[i]->getNumSolutions();
    if (!actions[i]->initialize(domain, graph, solutionRank % n)) return false;
    solutionRank /= n;
  }
  return true;
}

ActionDoubleGet::ActionDoubleGet(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
{
  if (params.size() != 2)
    throw ActionException("ERROR REASON: Action command received with " + std::to_string(params.size()) + " arguments, but 2 are expected",
                          ActionException::ParamInvalid);
  if (domain.manipulators.size() < 2)
    throw ActionException("ERROR REASON: Action created with " + std::to_string(domain.manipulators.size()) +
                          " manipulators, but 2 or more are expected", ActionException::ParamInvalid);
  const int obj1 = graph.getBodyByName(params[0]), obj2 = graph.getBodyByName(params[1]);
  if (obj1 < 0) throw ActionException("ERROR REASON: Object " + params[0] + " is unknown", ActionException::ParamNotFound);
  if (obj2 < 0) throw ActionException("ERROR REASON: Object " + params[1] + " is unknown", ActionException::ParamNotFound);

  // The pairing (hand for object 1, another hand for object 2) with the smallest summed distance
  std::vector<std::array<double, 2>> res;
  for (const Manipulator& m : domain.manipulators) {
    const int mBdy = graph.getBodyByName(m.name);
    if (mBdy < 0) throw ActionException("ERROR REASON: Manipulator " + m.name + " has no body", ActionException::UnknownError);
    auto dist = [&](int b) {
      double d2 = 0.0;
      for (int k = 0; k < 3; ++k) { const double d = graph.bodies[mBdy].A_BI.org[k] - graph.bodies[b].A_BI.org[k]; d2 += d * d; }
      return std::sqrt(d2);
    };
    res.push_back({dist(obj1), dist(obj2)});
  }
  double dMin = DBL_MAX;
  std::string m1, m2;
  for (size_t c1 = 0; c1 < res.size(); ++c1)
    for (size_t c2 = 0; c2 < res.size(); ++c2) {
      if (c1 == c2) continue;
      const double di = res[c1][0] + res[c2][1];
      if (di < dMin) { dMin = di; m1 = domain.manipulators[c1].name; m2 = domain.manipulators[c2].name; }
    }
  // The reference builds both sub-actions from params[0] (src/ActionComposite.cpp:240-243), which
  // makes both hands reach for the first object; its explanation ("Getting A and B") and the
  // pairing above show the second object is meant, and that is what is built here.
  addAction(std::unique_ptr<ActionBase>(new ActionGet(domain, graph, {params[0], m1})));
  addAction(std::unique_ptr<ActionBase>(new ActionGet(domain, graph, {params[1], m2})));
  explanation = "Getting " + params[0] + " and " + params[1];
}

ActionDoubleGetPairs::ActionDoubleGetPairs(const ActionScene& domain, const Graph& graph, std::vector<std::string> params)
{
  if (params.size() != 2)
    throw ActionException("ERROR REASON: Action command received with " + std::to_string(params.size()) + " arguments, but 2 are expected",
                          ActionException::ParamInvalid);
  addAction(std::unique_ptr<ActionBase>(new ActionGet(domain, graph, {params[0]})));
  addAction(std::unique_ptr<ActionBase>(new ActionGet(domain, graph, {params[1]})));
  const size_t n1 = actions[0]->getNumSolutions(), n2 = actions[1]->getNumSolutions();
  for (size_t i = 0; i < n1; ++i)
    for (size_t j = 0; j < n2; ++j) {
      if (!actions[0]->initialize(domain, graph, i) || !actions[1]->initialize(domain, graph, j)) continue;
      const std::vector<std::string> h1 = actions[0]->getManipulators(), h2 = actions[1]->getManipulators();
      if (!h1.empty() && !h2.empty() && h1[0] != h2[0]) pairs.emplace_back(i, j);
    }
  if (pairs.empty())
    throw ActionException("ERROR REASON: No pair of free hands can get " + params[0] + " and " + params[1], ActionException::KinematicallyImpossible);
  explanation = "Getting " + params[0] + " and " + params[1] + " (every hand pairing)";
}

bool ActionDoubleGetPairs::initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank)
{
  if (solutionRank >= pairs.size()) return false;
  return actions[0]->initialize(domain, graph, pairs[solutionRank].first) && actions[1]->initialize(domain, graph, pairs[solutionRank].second);
}

}   // namespace affgpu
