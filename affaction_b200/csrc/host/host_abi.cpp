// extern "C" surface of the host shim: scene loading, candidate generation
// from text commands, batch prediction with message strings.  This is what a
// non-C++ front end (the ctypes binding in affaction_b200/__init__.py, the
// pybind module of the reference's python/module.cpp) binds.
#include "gpu_api.h"
#include <cmath>
#include <cstring>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <algorithm>
#include <cstdlib>

#include "ActionBase.h"
#include "ActionComposite.h"
#include "ActionDrop.h"
#include "ActionFingerPush.h"
#include "ActionGet.h"
#include "ActionPose.h"
#include "ActionPour.h"
#include "ActionPut.h"
#include "TrajectoryComponent.h"
#include "TrajectoryPredictor.h"
#include "affpredict_host.h"

using namespace affgpu;

struct aff_host_scene
{
  Graph graph;
  ActionScene domain;
  aff_scene* device = nullptr;
  int deviceIndex = -1;          // the CUDA device `device` was created on
};

struct aff_host_batch
{
  CandidateBatch batch;
  std::vector<std::string> commands;   // per candidate
};

static thread_local std::string g_hostError;

extern "C" {

const char* aff_host_last_error(void) { return g_hostError.c_str(); }

aff_host_scene* aff_host_scene_load(const char* path)
{
  try {
    std::unique_ptr<aff_host_scene> s(new aff_host_scene);
    std::vector<std::string> tail;
    s->graph = Graph::load(path, &tail);
    s->domain.parse(tail);
    s->graph.toDesc();
    return s.release();
  } catch (const std::exception& e) {
    g_hostError = e.what();
    return nullptr;
  }
}

void aff_host_scene_destroy(aff_host_scene* s)
{
  if (!s) return;
  if (s->device) gpu().scene_destroy(s->device);
  delete s;
}

const aff_scene_desc* aff_host_scene_desc(aff_host_scene* s) { return s->graph.toDesc(); }
int aff_host_num_bodies(const aff_host_scene* s) { return (int)s->graph.bodies.size(); }
int aff_host_num_ik_joints(const aff_host_scene* s) { return s->graph.nJ; }
int aff_host_body_id(const aff_host_scene* s, const char* name) { return s->graph.getBodyByName(name); }
const char* aff_host_body_name(const aff_host_scene* s, int id)
{
  return (id >= 0 && id < (int)s->graph.bodies.size()) ? s->graph.bodies[id].name.c_str() : "NULL";
}
const char* aff_host_ik_joint_name(const aff_host_scene* s, int jacobiIndex)
{
  for (const Joint& j : s->graph.joints) if (j.jacobiIndex == jacobiIndex) return j.name.c_str();
  return "NULL";
}

int aff_host_scene_connect(aff_host_scene* s, const char* child, const char* parent)
{
  try {
    const int c = s->graph.getBodyByName(child), p = s->graph.getBodyByName(parent);
    if (c < 0 || p < 0) { g_hostError = "unknown body"; return AFF_ERR_INVALID; }
    s->graph.connectBody(c, p);
    if (s->graph.restoreParentFirstOrder()) s->graph.computeForwardKinematics();
    s->graph.toDesc();
    if (s->device) { gpu().scene_destroy(s->device); s->device = nullptr; }
    return AFF_OK;
  } catch (const std::exception& e) { g_hostError = e.what(); return AFF_ERR_INVALID; }
}

// Start state for the next action of a sequence: the joint values a finished rollout ended with
// (jacobi order, e.g. the last row of aff_batch_fetch_q).
int aff_host_scene_set_ik_state(aff_host_scene* s, const double* q, int n)
{
  if (!s || !q || n != s->graph.nJ) { g_hostError = "set_ik_state: wrong number of joint values"; return AFF_ERR_INVALID; }
  for (Joint& j : s->graph.joints) if (j.jacobiIndex >= 0) j.q_init = q[j.jacobiIndex];
  s->graph.computeForwardKinematics();
  s->graph.toDesc();
  if (s->device) { gpu().scene_destroy(s->device); s->device = nullptr; }
  return AFF_OK;
}

int aff_host_scene_apply_rollout(aff_host_scene* s, const aff_host_batch* b, size_t cand, const double* qlog, int rows, double dt)
{
  try {
    if (!s || !b || cand >= b->batch.candidates.size()) { g_hostError = "apply_rollout: candidate out of range"; return AFF_ERR_INVALID; }
    const aff_candidate& c = b->batch.candidates[cand];
    s->graph.applyRollout(b->batch.constraints.data() + c.first_constraint, c.n_constraints, qlog, rows, dt);
    s->graph.toDesc();
    if (s->device) { gpu().scene_destroy(s->device); s->device = nullptr; }
    return AFF_OK;
  } catch (const std::exception& e) { g_hostError = e.what(); return AFF_ERR_INVALID; }
}

int aff_host_rollout_transforms(aff_host_scene* s, const aff_host_batch* b, size_t cand, const double* qlog, int rows, double dt, double* out)
{
  try {
    if (!s || !b || cand >= b->batch.candidates.size()) { g_hostError = "rollout_transforms: candidate out of range"; return AFF_ERR_INVALID; }
    const aff_candidate& c = b->batch.candidates[cand];
    s->graph.rolloutTransforms(b->batch.constraints.data() + c.first_constraint, c.n_constraints, qlog, rows, dt, out);
    return AFF_OK;
  } catch (const std::exception& e) { g_hostError = e.what(); return AFF_ERR_INVALID; }
}

int aff_host_scene_desc_roundtrip(aff_host_scene* s)
{
  try {
    const aff_scene_desc* d = s->graph.toDesc();
    std::vector<std::string> bn, jn;
    for (const Body& b : s->graph.bodies) bn.push_back(b.name);
    for (const Joint& j : s->graph.joints) jn.push_back(j.name);
    Graph g2 = Graph::fromDesc(*d, bn, jn, s->graph.name);
    const aff_scene_desc* e = g2.toDesc();
    auto same = [](const void* a, const void* b, size_t n) { return n == 0 || std::memcmp(a, b, n) == 0; };
    const size_t nb = (size_t)d->n_bodies, nj = (size_t)d->n_joints, ns = (size_t)d->n_shapes;
    bool ok = d->n_bodies == e->n_bodies && d->n_joints == e->n_joints && d->n_shapes == e->n_shapes &&
              d->n_bp_trees == e->n_bp_trees && d->n_bp_bodies == e->n_bp_bodies && d->bp_threshold == e->bp_threshold;
    ok = ok && same(d->body_parent, e->body_parent, 4 * nb) && same(d->body_A_BP, e->body_A_BP, 96 * nb) && same(d->body_rbj, e->body_rbj, nb) &&
         same(d->body_first_joint, e->body_first_joint, 4 * nb) && same(d->body_first_shape, e->body_first_shape, 4 * nb) &&
         same(d->joint_type, e->joint_type, 4 * nj) && same(d->joint_A_JP, e->joint_A_JP, 96 * nj) && same(d->joint_q_init, e->joint_q_init, 8 * nj) &&
         same(d->joint_q_min, e->joint_q_min, 8 * nj) && same(d->joint_q_max, e->joint_q_max, 8 * nj) && same(d->joint_speed_limit, e->joint_speed_limit, 8 * nj) &&
         same(d->shape_type, e->shape_type, 4 * ns) && same(d->shape_A_CB, e->shape_A_CB, 96 * ns) && same(d->shape_extents, e->shape_extents, 24 * ns) &&
         same(d->bp_tree_root, e->bp_tree_root, 4 * (size_t)d->n_bp_trees) && same(d->bp_body, e->bp_body, 4 * (size_t)d->n_bp_bodies) &&
         d->ns_base == e->ns_base && d->ns_hand[0] == e->ns_hand[0] && d->ns_hand[1] == e->ns_hand[1];
    return ok ? 1 : 0;
  } catch (const std::exception& e) { g_hostError = e.what(); return AFF_ERR_INVALID; }
}

aff_scene* aff_host_scene_device(aff_host_scene* s, int device)
{
  // one cached device scene per host scene; asking for another device replaces it
  try {
    if (s->device && s->deviceIndex != device) { gpu().scene_destroy(s->device); s->device = nullptr; }
    if (!s->device) {
      if (gpu().scene_create(s->graph.toDesc(), device, &s->device) != AFF_OK) { g_hostError = gpu().last_error(); return nullptr; }
      s->deviceIndex = device;
    }
  } catch (const std::exception& e) { g_hostError = e.what(); return nullptr; }    // the CUDA library is missing
  return s->device;
}

aff_host_batch* aff_host_batch_create(void) { return new aff_host_batch; }
void aff_host_batch_destroy(aff_host_batch* b) { delete b; }
void aff_host_batch_clear(aff_host_batch* b) { b->batch = CandidateBatch(); b->commands.clear(); }

size_t aff_host_batch_num_tasks(const aff_host_batch* b) { return b->batch.tasks.size(); }
size_t aff_host_batch_num_constraints(const aff_host_batch* b) { return b->batch.constraints.size(); }
size_t aff_host_batch_num_candidates(const aff_host_batch* b) { return b->batch.candidates.size(); }
const aff_task_desc* aff_host_batch_tasks(const aff_host_batch* b) { return b->batch.tasks.data(); }
const aff_constraint* aff_host_batch_constraints(const aff_host_batch* b) { return b->batch.constraints.data(); }
const aff_candidate* aff_host_batch_candidates(const aff_host_batch* b) { return b->batch.candidates.data(); }
const char* aff_host_batch_task_name(const aff_host_batch* b, size_t cand, int task)
{
  if (cand >= b->batch.taskNames.size() || task < 0 || task >= (int)b->batch.taskNames[cand].size()) return "";
  return b->batch.taskNames[cand][task].c_str();
}

// The predict block of ActionComponent::actionThread (reference
// src/ActionComponent.cpp:145-199): create the action, then one candidate per
// solution: clone, initialize(i), describe.
int aff_host_batch_add_action(aff_host_batch* b, aff_host_scene* s, const char* text, double durationScale, double dt)
{
  try {
    auto action = ActionFactory::create(s->domain, s->graph, text);
    const size_t n = action->getNumSolutions();
    for (size_t i = 0; i < n; ++i) {
      auto local = action->clone();
      if (!local->initialize(s->domain, s->graph, i)) { g_hostError = "initialize failed"; return AFF_ERR_INVALID; }
      local->appendTo(b->batch, s->graph, durationScale * local->getDurationHint(), dt);
      b->commands.push_back(local->getActionCommand());
    }
    return (int)n;
  } catch (const std::exception& e) {
    g_hostError = e.what();
    return AFF_ERR_INVALID;
  }
}

// splitmix64: one independent stream per rollout index
static inline uint64_t splitmix(uint64_t& x)
{
  uint64_t z = (x += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
static inline double uni(uint64_t& st, double lo, double hi)
{
  return lo + (hi - lo) * ((double)(splitmix(st) >> 11) * (1.0 / 9007199254740992.0));
}

// Synthetic perturbed-goal sweep: rollout i uses hand
// i mod nSolutions, RNG seed seed0 + i, object position +U(-5,5) cm in x/y,
// +U(0,2) cm in z, yaw +U(-30,30) deg, preGraspDist x U(0.8,1.2), liftHeight
// +U(-3,3) cm.  The duration stays at the action's hint so every rollout runs
// the same, full number of steps.
int aff_host_batch_add_perturbed(aff_host_batch* b, aff_host_scene* s, const char* text, size_t n, uint64_t seed0, double dt)
{
  try {
    auto action = ActionFactory::create(s->domain, s->graph, text);
    ActionGet* get = dynamic_cast<ActionGet*>(action.get());
    if (!get) { g_hostError = "perturbed sweeps are defined for get actions"; return AFF_ERR_UNSUPPORTED; }
    const size_t nSol = action->getNumSolutions();
    std::vector<std::unique_ptr<ActionBase>> sols;
    for (size_t r = 0; r < nSol; ++r) { sols.push_back(action->clone()); sols.back()->initialize(s->domain, s->graph, r); }
    const int obj = s->graph.getBodyByName(get->objectName);
    const Body& ob = s->graph.bodies[obj];
    if (!ob.rigid_body_joints) { g_hostError = "object has no rigid body joints"; return AFF_ERR_INVALID; }
    double q6[6];
    for (int k = 0; k < 6; ++k) q6[k] = s->graph.joints[ob.firstJoint + k].q_init;
    // candidate i depends on (seed0 + i, i mod nSol) only: slices of the index range are described by host
    // threads into private batches and appended in index order (AFF_HOST_THREADS, default: the hardware
    // threads, at most 16; below 2048 candidates one thread)
    // the task list of a command and hand does not depend on the perturbation: one prototype per solution
    // (AFF_HOST_PERTURBED_GENERIC=1: every candidate through the generic append(), for the equality test)
    CandidateBatch proto;
    const bool generic = getenv("AFF_HOST_PERTURBED_GENERIC") != nullptr;
    for (size_t r = 0; r < nSol; ++r) {
      const ActionGet& g = *static_cast<ActionGet*>(sols[r].get());
      proto.append(s->graph, g.createTasks(), ActivationSet());
    }
    auto describe = [&](size_t lo, size_t hi, CandidateBatch& out, std::vector<std::string>& cmds) {
      out.candidates.reserve(out.candidates.size() + (hi - lo));
      for (size_t i = lo; i < hi; ++i) {
        uint64_t st = seed0 + i;
        ActionGet local(*static_cast<ActionGet*>(sols[i % nSol].get()));
        double p6[6];
        for (int k = 0; k < 6; ++k) p6[k] = q6[k];
        p6[0] += uni(st, -0.05, 0.05);
        p6[1] += uni(st, -0.05, 0.05);
        const double dz = uni(st, 0.0, 0.02);
        p6[2] += dz;
        p6[5] += uni(st, -30.0, 30.0) * (M_PI / 180.0);
        local.preGraspDist *= uni(st, 0.8, 1.2);
        local.liftHeight += dz + uni(st, -0.03, 0.03);
        const double delay = 5.0 * dt;
        if (generic) out.append(s->graph, local.createTasks(), local.createTrajectory(delay, local.getDurationHint() + delay), obj, p6);
        else out.appendLike(s->graph, proto, i % nSol, local.createTrajectory(delay, local.getDurationHint() + delay), obj, p6);
        cmds.push_back(local.getActionCommand());
      }
    };
    size_t nThreads = 1;
    if (n >= 2048) {
      const char* env = getenv("AFF_HOST_THREADS");
      nThreads = env ? (size_t)std::max(1, atoi(env)) : std::min<size_t>(16, std::max(1u, std::thread::hardware_concurrency()));
      nThreads = std::min(nThreads, n / 1024);
    }
    if (nThreads <= 1) { describe(0, n, b->batch, b->commands); return (int)n; }
    std::vector<CandidateBatch> part(nThreads);
    std::vector<std::vector<std::string>> partCmds(nThreads);
    std::vector<std::string> errs(nThreads);
    std::vector<std::thread> pool;
    for (size_t t = 0; t < nThreads; ++t)
      pool.emplace_back([&, t]() {
        try { describe(n * t / nThreads, n * (t + 1) / nThreads, part[t], partCmds[t]); }
        catch (const std::exception& e) { errs[t] = e.what(); if (errs[t].empty()) errs[t] = "error"; }
      });
    for (std::thread& th : pool) th.join();
    for (size_t t = 0; t < nThreads; ++t) if (!errs[t].empty()) { g_hostError = errs[t]; return AFF_ERR_INVALID; }
    CandidateBatch& B = b->batch;
    size_t addT = 0, addK = 0;
    for (const CandidateBatch& p : part) { addT += p.tasks.size(); addK += p.constraints.size(); }
    B.tasks.reserve(B.tasks.size() + addT); B.constraints.reserve(B.constraints.size() + addK);
    B.candidates.reserve(B.candidates.size() + n); B.taskNames.reserve(B.taskNames.size() + n); b->commands.reserve(b->commands.size() + n);
    for (size_t t = 0; t < nThreads; ++t) {
      const int32_t t0 = (int32_t)B.tasks.size(), k0 = (int32_t)B.constraints.size();
      B.tasks.insert(B.tasks.end(), part[t].tasks.begin(), part[t].tasks.end());
      B.constraints.insert(B.constraints.end(), part[t].constraints.begin(), part[t].constraints.end());
      for (aff_candidate c : part[t].candidates) { c.first_task += t0; c.first_constraint += k0; B.candidates.push_back(c); }
      for (auto& nm : part[t].taskNames) B.taskNames.push_back(std::move(nm));
      for (auto& cm : partCmds[t]) b->commands.push_back(std::move(cm));
      part[t] = CandidateBatch();
    }
    return (int)n;
  } catch (const std::exception& e) {
    g_hostError = e.what();
    return AFF_ERR_INVALID;
  }
}

// Predict all candidates on the GPU, optionally sorted with
// PredictionResult::lesser; messages are written as NUL-terminated strings of
// at most msgStride-1 characters each.
int aff_host_predict(aff_host_scene* s, aff_host_batch* b, double dt, int sortResults, int device,
                     aff_result* rawOut, char* messages, size_t msgStride)
{
  try {
    aff_scene* dev = aff_host_scene_device(s, device);
    if (!dev) return AFF_ERR_CUDA;
    auto res = TrajectoryPredictor::predictBatch(dev, s->graph, b->batch, dt, sortResults != 0);
    for (size_t i = 0; i < res.size(); ++i) {
      if (rawOut) rawOut[i] = res[i].raw;
      if (messages && msgStride > 0) {
        std::string m = res[i].message;
        // multi-candidate path appends the command (reference src/ActionComponent.cpp:193)
        if (res[i].idx >= 0 && (size_t)res[i].idx < b->commands.size()) m += " command: " + b->commands[res[i].idx];
        std::strncpy(messages + i * msgStride, m.c_str(), msgStride - 1);
        messages[i * msgStride + msgStride - 1] = '\0';
      }
    }
    return (int)res.size();
  } catch (const std::exception& e) {
    g_hostError = e.what();
    return AFF_ERR_CUDA;
  }
}

// Message for a raw record (used by tests to compare oracle and GPU wording).
int aff_host_format_message(aff_host_scene* s, aff_host_batch* b, const aff_result* r, double dt, char* out, size_t n)
{
  static const std::vector<std::string> none;
  const auto& names = (r->idx >= 0 && (size_t)r->idx < b->batch.taskNames.size()) ? b->batch.taskNames[r->idx] : none;
  const auto p = TrajectoryPredictor::convert(*r, s->graph, names, dt);
  std::strncpy(out, p.message.c_str(), n - 1);
  out[n - 1] = '\0';
  return (int)p.message.size();
}

int aff_host_format_message_q(aff_host_scene* s, aff_host_batch* b, const aff_result* r, double dt, const double* qFail, char* out, size_t n)
{
  static const std::vector<std::string> none;
  const auto& names = (r->idx >= 0 && (size_t)r->idx < b->batch.taskNames.size()) ? b->batch.taskNames[r->idx] : none;
  const auto p = TrajectoryPredictor::convert(*r, s->graph, names, dt, qFail);
  std::strncpy(out, p.message.c_str(), n - 1);
  out[n - 1] = '\0';
  return (int)p.message.size();
}

// "a; b; c": predict, take the ranked-first candidate, apply its end state, go on.
int aff_host_predict_sequence(aff_host_scene* s, const char* text, double dt, int device,
                              aff_result* winners, int* nCandidates, char* messages, size_t msgStride, int maxActions)
{
  try {
    std::vector<std::string> actions;
    {
      std::istringstream is(text ? text : "");
      for (std::string a; std::getline(is, a, ';');) {
        const size_t b0 = a.find_first_not_of(" \t\n"), b1 = a.find_last_not_of(" \t\n");
        if (b0 != std::string::npos) actions.push_back(a.substr(b0, b1 - b0 + 1));
      }
    }
    int done = 0;
    for (const std::string& text_i : actions) {
      if (done >= maxActions) break;
      CandidateBatch batch;
      std::vector<std::string> commands;
      auto action = ActionFactory::create(s->domain, s->graph, text_i);
      const size_t n = action->getNumSolutions();
      for (size_t i = 0; i < n; ++i) {
        auto local = action->clone();
        if (!local->initialize(s->domain, s->graph, i)) { g_hostError = "initialize failed"; return AFF_ERR_INVALID; }
        local->appendTo(batch, s->graph, local->getDurationHint(), dt);
      }
      aff_scene* dev = aff_host_scene_device(s, device);
      if (!dev) return AFF_ERR_CUDA;
      aff_opts opts;
      aff_default_opts(&opts);
      opts.dt = dt;
      opts.log_q = 1;                 // the winner's joint trajectory is the next action's start
      aff_batch* db = nullptr;
      if (gpu().batch_create(dev, batch.tasks.data(), batch.tasks.size(), batch.constraints.data(), batch.constraints.size(),
                           batch.candidates.data(), batch.candidates.size(), &opts, &db) != AFF_OK) { g_hostError = gpu().last_error(); return AFF_ERR_CUDA; }
      std::vector<aff_result> raw(batch.candidates.size());
      int rc = gpu().batch_launch(db, nullptr);
      if (rc == AFF_OK) rc = gpu().batch_results(db, nullptr, raw.data(), raw.size());
      std::vector<int32_t> order(raw.size());
      aff_rank_results(raw.data(), raw.size(), order.data());
      const aff_result best = raw[order[0]];
      std::vector<double> q;
      int rows = 0;
      if (rc == AFF_OK && best.success) {
        rows = gpu().batch_num_steps(db, (size_t)best.idx) + 1;
        q.resize((size_t)rows * s->graph.nJ);
        rows = gpu().batch_fetch_q(db, (size_t)best.idx, q.data(), (size_t)rows);
        if (rows < 0) rc = rows;
      }
      gpu().batch_destroy(db);
      if (rc != AFF_OK) { g_hostError = gpu().last_error(); return AFF_ERR_CUDA; }
      if (winners) winners[done] = best;
      if (nCandidates) nCandidates[done] = (int)raw.size();
      if (messages && msgStride > 0) {
        const std::string m = TrajectoryPredictor::convert(best, s->graph, batch.taskNames[best.idx], dt).message + " command: " + text_i;
        std::strncpy(messages + (size_t)done * msgStride, m.c_str(), msgStride - 1);
        messages[(size_t)done * msgStride + msgStride - 1] = '\0';
      }
      ++done;
      if (!best.success) break;         // the reference clears the rest of the queue when an action fails
      const aff_candidate& c = batch.candidates[best.idx];
      s->graph.applyRollout(batch.constraints.data() + c.first_constraint, c.n_constraints, q.data(), rows, dt);
      s->graph.toDesc();
      if (s->device) { gpu().scene_destroy(s->device); s->device = nullptr; }
    }
    return done;
  } catch (const std::exception& e) {
    g_hostError = e.what();
    return AFF_ERR_INVALID;
  }
}

// TrajectoryComponent on one candidate of a batch: its task list plays the runtime controller's,
// its constraints the trajectory to check.
int aff_host_check_trajectory(aff_host_scene* s, aff_host_batch* b, size_t cand, int mode, int estop, int checkEnabled,
                              double dt, int device, int* actionResult, int* ok, double* quality, int* setTrajectory,
                              aff_result* rawOut, char* message, size_t msgLen, double* qOut, size_t maxRows)
{
  try {
    if (cand >= b->batch.candidates.size()) { g_hostError = "check_trajectory: candidate out of range"; return AFF_ERR_INVALID; }
    aff_scene* dev = aff_host_scene_device(s, device);
    if (!dev) return AFF_ERR_CUDA;
    // rebuild TaskSpec-free: the component works on descriptors already resolved against the graph
    const aff_candidate& c = b->batch.candidates[cand];
    CandidateBatch one;
    one.tasks.assign(b->batch.tasks.begin() + c.first_task, b->batch.tasks.begin() + c.first_task + c.n_tasks);
    one.constraints.assign(b->batch.constraints.begin() + c.first_constraint, b->batch.constraints.begin() + c.first_constraint + c.n_constraints);
    aff_candidate c1 = c;
    c1.first_task = 0; c1.first_constraint = 0;
    one.candidates.push_back(c1);
    one.taskNames.push_back(b->batch.taskNames[cand]);
    TrajectoryComponent comp(dev, s->graph, dt);
    if (estop) comp.onEmergencyStop();
    comp.onEnableTrajectoryCheck(checkEnabled != 0);
    const TrajectoryComponent::CheckOutcome out = mode == 1 ? comp.onSimulateTrajectory(one) : comp.onCheckAndSetTrajectory(one);
    if (actionResult) *actionResult = out.publishActionResult ? 1 : 0;
    if (ok) *ok = out.actionOk ? 1 : 0;
    if (quality) *quality = out.quality;
    if (setTrajectory) *setTrajectory = out.publishSetTrajectory ? 1 : 0;
    if (rawOut && out.checked) *rawOut = out.result.raw;
    if (message && msgLen > 0) { std::strncpy(message, out.message.c_str(), msgLen - 1); message[msgLen - 1] = '\0'; }
    int rows = 0, cols = 0;
    const std::vector<double>& q = comp.getPredictionArray(&rows, &cols);
    if (qOut) std::memcpy(qOut, q.data(), sizeof(double) * std::min((size_t)rows, maxRows) * cols);
    return rows;
  } catch (const std::exception& e) {
    g_hostError = e.what();
    return AFF_ERR_CUDA;
  }
}

}   // extern "C"

namespace affgpu
{

std::unique_ptr<ActionBase> ActionFactory::create(const ActionScene& domain, const Graph& graph, const std::string& text)
{
  std::istringstream is(text);
  std::vector<std::string> words;
  for (std::string w; is >> w;) words.push_back(w);
  if (words.empty()) throw ActionException("ERROR: empty action command", ActionException::ParamNotFound);
  const std::string name = words[0];
  std::vector<std::string> params(words.begin() + 1, words.end());
  std::unique_ptr<ActionBase> a;
  if (name == "get") a.reset(new ActionGet(domain, graph, params));
  else if (name == "put") a.reset(new ActionPut(domain, graph, params));
  else if (name == "double_get") a.reset(new ActionDoubleGet(domain, graph, params));
  else if (name == "double_get_pairs") a.reset(new ActionDoubleGetPairs(domain, graph, params));
  else if (name == "drop") a.reset(new ActionDrop(domain, graph, params));
  else if (name == "pose") a.reset(new ActionPose(domain, graph, params));
  else if (name == "finger_push" || name == "poke") a.reset(new ActionFingerPush(domain, graph, params));
  else if (name == "pour") a.reset(new ActionPour(domain, graph, params));
  else
    throw ActionException("ERROR: REASON: Action '" + name + "' is not part of the GPU predict path" +
                          " SUGGESTION: Use one of: get, put, pour, drop, pose, finger_push, double_get, double_get_pairs", ActionException::ParamNotFound);
  a->setName(name);
  a->setActionParams(params);
  return a;
}

}   // namespace affgpu
