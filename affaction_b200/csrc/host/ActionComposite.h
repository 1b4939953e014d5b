// Composite actions: several sub-actions run in the same rollout, their task lists unioned and
// their trajectories merged (reference src/ActionComposite.cpp:60-150), and the one composite the
// predict path is benchmarked on, "double_get <object 1> <object 2>" (:158-262).
#ifndef AFF_HOST_ACTIONCOMPOSITE_H
#define AFF_HOST_ACTIONCOMPOSITE_H

#include "ActionBase.h"

namespace affgpu
{

class ActionComposite : public ActionBase
{
public:
  ActionComposite() {}
  ActionComposite(const ActionComposite& other);
  void addAction(std::unique_ptr<ActionBase> action) { actions.push_back(std::move(action)); }

  std::vector<TaskSpec> createTasks() const override;
  ActivationSet createTrajectory(double t_start, double t_end) const override;
  using ActionBase::createTrajectory;
  double getDurationHint() const override;
  std::vector<std::string> getManipulators() const override;

  // Extension: the reference leaves every sub-action at its first solution,
  // i.e. one candidate.  Here the candidates are the Cartesian product of the sub-actions'
  // solutions, rank = mixed-radix index with the last sub-action fastest; rank 0 is the reference's.
  size_t getNumSolutions() const override;
  bool initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank) override;

protected:
  std::vector<std::unique_ptr<ActionBase>> actions;
};

class ActionDoubleGet : public ActionComposite
{
public:
  ActionDoubleGet(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  ActionDoubleGet(const ActionDoubleGet& other) : ActionComposite(other), explanation(other.explanation) {}
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionDoubleGet(*this)); }
  std::string explain() const override { return explanation; }

private:
  std::string explanation;
};

// Extension named in SURVEY 8(d) config 3 / BASELINE.json configs[2] ("bimanual candidate pairs"):
// "double_get_pairs <object 1> <object 2>".  The candidate set is the Cartesian product
// affordanceMap(object 1) x affordanceMap(object 2) of the two gets' Capability x Affordance solutions,
// minus the pairs that would use one hand twice; candidate order = the product's mixed-radix order.
// The reference's double_get (above) pins one hand pairing and so yields exactly one candidate.
class ActionDoubleGetPairs : public ActionComposite
{
public:
  ActionDoubleGetPairs(const ActionScene& domain, const Graph& graph, std::vector<std::string> params);
  ActionDoubleGetPairs(const ActionDoubleGetPairs& other) : ActionComposite(other), pairs(other.pairs), explanation(other.explanation) {}
  std::unique_ptr<ActionBase> clone() const override { return std::unique_ptr<ActionBase>(new ActionDoubleGetPairs(*this)); }
  std::string explain() const override { return explanation; }
  size_t getNumSolutions() const override { return pairs.size(); }
  bool initialize(const ActionScene& domain, const Graph& graph, size_t solutionRank) override;

private:
  std::vector<std::pair<size_t, size_t>> pairs;     // (solution of get 1, solution of get 2), hands distinct
  std::string explanation;
};

}   // namespace affgpu

#endif
