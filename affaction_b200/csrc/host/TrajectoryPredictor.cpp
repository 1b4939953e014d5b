#include "gpu_api.h"
#include "TrajectoryPredictor.h"

#include <algorithm>
#include <stdexcept>

#include "ActionBase.h"

namespace affgpu
{

static std::string bodyName(const Graph& g, int id)
{
  return (id >= 0 && id < (int)g.bodies.size()) ? g.bodies[id].name : std::string("NULL");
}

// RcsBody_isArticulated: some unconstrained joint moves the body.
static bool isArticulated(const Graph& g, int body)
{
  for (int b = body; b >= 0; b = g.bodies[b].parent)
    for (int k = 0; k < g.bodies[b].nJoints; ++k)
      if (!g.joints[g.bodies[b].firstJoint + k].constrained) return true;
  return false;
}

TrajectoryPredictor::PredictionResult TrajectoryPredictor::convert(const aff_result& r, const Graph& graph,
                                                                   const std::vector<std::string>& taskNames, double dt, const double* qFail)
{
  PredictionResult p;
  p.raw = r;
  p.idx = r.idx;
  p.success = r.success != 0;
  p.minDist = r.min_dist; p.jlCost = r.jl_cost; p.collCost = r.coll_cost;
  p.minDistBdy1 = bodyName(graph, r.min_dist_b1);
  p.minDistBdy2 = bodyName(graph, r.min_dist_b2);
  p.jMask.assign((size_t)graph.nJ, 0.0);
  for (int c = 0; c < graph.nJ && c < 64; ++c) if (r.jmask_bits & (1ull << c)) p.jMask[c] = 1.0;

  switch (r.code) {
    case AFF_CODE_OK:
      p.message = "SUCCESS";
      break;
    case AFF_CODE_INITIAL_STATE:
      p.message = "FATAL_ERROR: Initial state is invalid";
      break;
    case AFF_CODE_SINGULAR:
      p.message = "ERROR: REASON: Got into a singular posture";
      break;
    case AFF_CODE_SPEED:
      p.message = "ERROR: Speed limit violated, I can't move that fast + REASON: The duration of the action might have "
                  "been too short. SUGGESTION: Call the action with a longer duration.";
      break;
    case AFF_CODE_JOINT_LIMIT:
      p.message = "ERROR: REASON: Joint limit problem. The robot is not able to do move well enough. ";
      p.message += " DEVELOPER: " + std::to_string(r.fail_id0) + " joint limit violations";
      if (qFail) {
        for (const Joint& j : graph.joints) {
          if (j.constrained || j.jacobiIndex < 0) continue;
          const double qi = qFail[j.jacobiIndex];
          if (qi < j.q_min || qi > j.q_max)
            p.message += " Joint " + j.name + ": " + std::to_string(qi) + " outside [" + std::to_string(j.q_min) + " ... " + std::to_string(j.q_max) + "]";
        }
      }
      break;
    case AFF_CODE_COLLISION: {
      const std::string b1 = bodyName(graph, r.fail_id0), b2 = bodyName(graph, r.fail_id1);
      const bool a1 = r.fail_id0 >= 0 && isArticulated(graph, r.fail_id0);
      const bool a2 = r.fail_id1 >= 0 && isArticulated(graph, r.fail_id1);
      if (a1) p.message = a2 ? "ERROR: REASON: Collision problem, " + b1 + " and " + b2 + " collide"
                             : "ERROR: REASON: Collision problem, " + b1 + " collides with the " + b2;
      else p.message = "ERROR: REASON: Collision problem, " + b2 + " collides with the " + b1;
      p.message += " SUGGESTION: Choose another command or try to remove an obstacle.";
      p.message += " DEVELOPER: Collision detected between " + b1 + " and " + b2 + " at step " + std::to_string(r.fail_step);
      break;
    }
    case AFF_CODE_TRACKING: {
      double t = 0.0;
      for (int i = 0; i < r.fail_step; ++i) t += dt;   // the reference accumulates t += dt
      p.message = "ERROR: Reachability problem REASON: Can't reach the object - it is too far away";
      p.message += " SUGGESTION: Try with the other hand or get it closer with a tool";
      p.message += " DEVELOPER: Tracking error at t=" + std::to_string(t) + " is " + std::to_string(r.track_err);
      if (r.fail_id0 >= 0 && r.fail_id0 < (int)taskNames.size())
        p.message += ": Task " + taskNames[r.fail_id0] + "[dim " + std::to_string(r.fail_id1) + "]\n";
      break;
    }
    default:
      p.message = "ERROR: unknown predictor code " + std::to_string(r.code);
  }
  if (r.reserved & AFF_FLAG_LIST_OVERFLOW)
    p.message += " DEVELOPER: a collision scratch list overflowed during this rollout (more than 32 pairs in contact or 64 AABB-found pairs in one step); collision cost / minimum distance may be incomplete";
  return p;
}

std::vector<TrajectoryPredictor::PredictionResult> TrajectoryPredictor::predictBatch(aff_scene* scene, const Graph& graph,
                                                                                     const CandidateBatch& batch, double dt,
                                                                                     bool sortResults)
{
  aff_opts opts;
  aff_default_opts(&opts);
  opts.dt = dt;
  std::vector<aff_result> raw(batch.candidates.size());
  const int rc = gpu().predict_batch(scene, batch.tasks.data(), batch.tasks.size(), batch.constraints.data(),
                                   batch.constraints.size(), batch.candidates.data(), batch.candidates.size(), &opts, raw.data());
  if (rc != AFF_OK) throw std::runtime_error(std::string("aff_predict_batch failed: ") + gpu().last_error());
  std::vector<PredictionResult> res;
  res.reserve(raw.size());
  for (size_t i = 0; i < raw.size(); ++i) res.push_back(convert(raw[i], graph, batch.taskNames[i], dt));
  if (sortResults) std::stable_sort(res.begin(), res.end(), PredictionResult::lesser);   // src/ActionComponent.cpp:244
  return res;
}

}   // namespace affgpu
