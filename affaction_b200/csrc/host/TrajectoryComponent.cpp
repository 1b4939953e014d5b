#include "gpu_api.h"
#include "TrajectoryComponent.h"

#include <stdexcept>

namespace affgpu
{

TrajectoryComponent::TrajectoryComponent(aff_scene* scene_, const Graph& graph_, double dt_)
  : scene(scene_), graph(graph_), dt(dt_), eStop(false), enableTrajectoryCheck(true), tPredRows(0), tPredCols(0)
{
}

TrajectoryComponent::CheckOutcome TrajectoryComponent::onCheckAndSetTrajectory(const CandidateBatch& one)
{
  CheckOutcome out;
  if (eStop) {
    out.publishActionResult = true;
    out.actionOk = false;
    out.message = "FATAL_ERROR REASON: Emergency-Stop triggered";
    return out;
  }
  if (!enableTrajectoryCheck) {
    out.publishSetTrajectory = true;     // passed on unchecked
    return out;
  }
  return checkerThread(one, false);
}

TrajectoryComponent::CheckOutcome TrajectoryComponent::checkerThread(const CandidateBatch& batch, bool simulateOnly)
{
  std::lock_guard<std::mutex> lock(checkerThreadMtx);
  CheckOutcome out;
  if (batch.candidates.size() != 1) throw std::runtime_error("TrajectoryComponent: exactly one trajectory is checked per call");

  aff_opts opts;
  aff_default_opts(&opts);
  opts.dt = dt;
  opts.log_q = 1;
  aff_batch* b = nullptr;
  if (gpu().batch_create(scene, batch.tasks.data(), batch.tasks.size(), batch.constraints.data(), batch.constraints.size(),
                       batch.candidates.data(), 1, &opts, &b) != AFF_OK)
    throw std::runtime_error(std::string("TrajectoryComponent: ") + gpu().last_error());
  aff_result raw;
  int rc = gpu().batch_launch(b, nullptr);
  if (rc == AFF_OK) rc = gpu().batch_results(b, nullptr, &raw, 1);
  if (rc == AFF_OK) {
    const int rows = gpu().batch_num_steps(b, 0) + 1, cols = gpu().scene_num_ik_joints(scene);
    std::vector<double> newPredictions((size_t)rows * cols);
    const int got = gpu().batch_fetch_q(b, 0, newPredictions.data(), (size_t)rows);
    if (got < 0) rc = got;
    else { newPredictions.resize((size_t)got * cols); tPred.swap(newPredictions); tPredRows = got; tPredCols = cols; }
  }
  gpu().batch_destroy(b);
  if (rc != AFF_OK) throw std::runtime_error(std::string("TrajectoryComponent: ") + gpu().last_error());

  out.checked = true;
  out.result = TrajectoryPredictor::convert(raw, graph, batch.taskNames[0], dt);
  if (!out.result.success) {
    out.publishActionResult = true;
    out.actionOk = false;
    out.message = out.result.message;
    return out;
  }
  if (simulateOnly) {
    out.publishActionResult = true;
    out.actionOk = true;
    out.quality = 1.0 / (1.0 + out.result.jlCost);
    out.message = "SUCCESS DEVELOPER: Simulated trajectory is valid";
  } else {
    out.publishSetTrajectory = true;
  }
  return out;
}

}   // namespace affgpu
