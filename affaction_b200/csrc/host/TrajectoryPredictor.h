// Host shim over the C ABI with the reference's TrajectoryPredictor surface
// (src/TrajectoryPredictor.h:48-157): PredictionResult, its ordering, and a
// predictor that takes a task list + constraint set and returns a result.
// The forward simulation itself runs on the GPU (affpredict.h); there is no
// CPU path behind these calls.
#ifndef AFF_HOST_TRAJECTORYPREDICTOR_H
#define AFF_HOST_TRAJECTORYPREDICTOR_H

#include <string>
#include <vector>

#include "Graph.h"
#include "affpredict.h"

namespace affgpu
{

struct CandidateBatch;

class TrajectoryPredictor
{
public:
  struct PredictionResult
  {
    PredictionResult() : idx(-1), success(false), minDist(0.0), jlCost(0.0), collCost(0.0), elbowNS(0.0), wristNS(0.0) {}

    double quality() const { return jlCost + collCost; }

    // success first, then lower jlCost + collCost (src/TrajectoryPredictor.h:61-74)
    static bool lesser(const PredictionResult& a, const PredictionResult& b)
    {
      if (a.success != b.success) return a.success;
      return a.quality() < b.quality();
    }

    int idx;
    bool success;
    double minDist, jlCost, collCost;
    double elbowNS, wristNS;               // never written by the reference either (:600,605)
    std::string message;
    std::string minDistBdy1, minDistBdy2;
    std::vector<double> jMask;             // nJ entries, 0 / 1
    std::vector<double> optimizationParameters;
    std::vector<double> bodyTransforms;    // opt-in only (68 MB per candidate): Graph::rolloutTransforms / aff_host_rollout_transforms
    aff_result raw;                        // the record the GPU wrote
  };

  // Predicts every candidate of `batch` in one GPU launch and converts the
  // records to PredictionResults with the reference's message strings.
  // Replaces the ConcurrentExecutor loop of src/ActionComponent.cpp:165-206.
  static std::vector<PredictionResult> predictBatch(aff_scene* scene, const Graph& graph, const CandidateBatch& batch,
                                                    double dt, bool sortResults = false);

  // aff_result -> PredictionResult, rebuilding `message` as checkState() /
  // predict() word it (src/TrajectoryPredictor.cpp:137,309-324,770,807-818,844-867).
  // qFail (optional): joint values of the failing step, jacobi order -> per-joint detail of a joint-limit message (:810-818)
  static PredictionResult convert(const aff_result& r, const Graph& graph, const std::vector<std::string>& taskNames, double dt,
                                  const double* qFail = nullptr);
};

}   // namespace affgpu

#endif
