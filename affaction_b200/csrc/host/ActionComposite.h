// Composite actions: several sub-actions run in the same rollout, their task lists unioned and
// their trajectories merged (reference src/ActionComposite.cpp:60-150), and the one composite the
// predict path is benchmarked on, "double_get <object 1> <object 2>" (:158-262).
#ifndef AFF_HOST_ACTIONCOMPOSITE_H
#define AFF_HOST_ACTIONCOMPOSITE_H

#include "ActionBase.h"

namespace aff
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

}   // namespace aff

#endif
