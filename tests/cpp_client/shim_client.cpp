// A C++ client of the host shim as INTEGRATION.md section 2 uses it inside the reference's
// ActionComponent::actionThread - compiled and linked against libaffhost.so by tests/test_abi.py.
// The dummy classes in `namespace aff` carry the names of the reference's own classes
// (src/ActionBase.h:50, src/ActionScene.h, src/TrajectoryPredictor.h): the shim lives in `namespace affgpu`,
// so both sets of names coexist in one binary ("What's weak" #8: ODR clash).
#include <cstdio>
#include <memory>
#include <string>
#include <vector>

#include "ActionBase.h"            // affgpu::ActionBase, affgpu::ActionFactory, affgpu::CandidateBatch
#include "ActionScene.h"           // affgpu::ActionScene
#include "Graph.h"                 // affgpu::Graph
#include "TrajectoryPredictor.h"   // affgpu::TrajectoryPredictor
#include "affpredict.h"

namespace aff
{
class ActionScene { public: int entities = 0; };
class ActionBase { public: virtual ~ActionBase() {} virtual size_t getNumSolutions() const { return 0; } };
class TrajectoryPredictor { public: struct PredictionResult { int idx = -1; bool success = false; double jlCost = 0, collCost = 0; std::string message; }; };
}   // namespace aff

int main(int argc, char** argv)
{
  if (argc < 3) { std::fprintf(stderr, "usage: %s scene.affscene \"command\" [--predict]\n", argv[0]); return 2; }
  const bool predict = argc > 3 && std::string(argv[3]) == "--predict";
  aff::ActionScene referenceDomain;                      // the reference's objects stay what they are
  (void)referenceDomain;
  std::vector<std::string> tail;
  affgpu::Graph graph = affgpu::Graph::load(argv[1], &tail);
  affgpu::ActionScene domain;
  domain.parse(tail);
  // description -> graph -> description (what integration/rcs_graph_adapter.h + Graph::fromDesc do in the reference tree)
  std::vector<std::string> bn, jn;
  for (const auto& b : graph.bodies) bn.push_back(b.name);
  for (const auto& j : graph.joints) jn.push_back(j.name);
  affgpu::Graph graph2 = affgpu::Graph::fromDesc(*graph.toDesc(), bn, jn);
  if (graph2.bodies.size() != graph.bodies.size() || graph2.nJ != graph.nJ) { std::fprintf(stderr, "fromDesc lost something\n"); return 1; }

  const double dt = 0.01;
  std::unique_ptr<affgpu::ActionBase> action = affgpu::ActionFactory::create(domain, graph, argv[2]);
  affgpu::CandidateBatch batch;
  for (size_t i = 0; i < action->getNumSolutions(); ++i) {
    auto local = action->clone();
    if (!local->initialize(domain, graph, i)) return 1;
    local->appendTo(batch, graph, local->getDurationHint(), dt);
  }
  std::printf("solutions %zu tasks %zu constraints %zu\n", action->getNumSolutions(), batch.tasks.size(), batch.constraints.size());
  if (!predict) return 0;

  aff_scene* gpuScene = nullptr;
  if (aff_scene_create(graph.toDesc(), 0, &gpuScene) != AFF_OK) { std::fprintf(stderr, "no device: %s\n", aff_last_error()); return 1; }
  std::vector<aff::TrajectoryPredictor::PredictionResult> predResults;        // the reference's result vector
  for (const auto& r : affgpu::TrajectoryPredictor::predictBatch(gpuScene, graph, batch, dt, /*sort=*/true)) {
    aff::TrajectoryPredictor::PredictionResult p;
    p.idx = r.idx; p.success = r.success; p.jlCost = r.jlCost; p.collCost = r.collCost; p.message = r.message;
    predResults.push_back(p);
    std::printf("%d %d %.17g | %.60s\n", p.idx, (int)p.success, p.jlCost + p.collCost, p.message.c_str());
  }
  aff_scene_destroy(gpuScene);
  return 0;
}
