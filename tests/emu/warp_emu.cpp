// TEST-ONLY 32-lane warp emulator.  NEVER linked into libaffpredict.so and
// never used by the product: it exists so the kernel source
// (affaction_b200/csrc/kernel/rollout.cuh) can be stepped on a CPU next to the
// oracle when debugging, and so the CPU test tier can exercise the plan
// compiler + kernel logic without a GPU.  Results obtained through it are not
// GPU results and are never reported as such.
//
// Each lane is a ucontext fiber; w_sync() yields to a round-robin scheduler, so
// a barrier completes when all 32 lanes have reached it.  Shuffles and ballots
// exchange values through a 32-entry array between two barriers.
#define AFF_EMULATE 1
#include <ucontext.h>

#include <cstdio>
#include <stdio.h>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <string>
#include <vector>

#include "kernel/plan_build.h"

// up to EMU_WARPS warps of one block: fiber f = 32 * warp + lane.  Every fiber advances to its next
// yield once per scheduler round, so the lanes of a warp stay aligned at their barriers, and flags
// written by another warp are seen a round later at the latest.
#define EMU_WARPS 8
static ucontext_t g_main, g_ctx[32 * EMU_WARPS];
static int g_cur = 0;
static bool g_done[32 * EMU_WARPS];
static std::function<void()>* g_body = nullptr;
static double g_xd[32 * EMU_WARPS];
static int g_xi[32 * EMU_WARPS];
static std::vector<char> g_stacks;

int w_lane() { return g_cur & 31; }
static int emu_warp() { return g_cur >> 5; }
void w_sync() { swapcontext(&g_ctx[g_cur], &g_main); }
double w_shfl(double v, int src) { g_xd[g_cur] = v; w_sync(); const double r = g_xd[(g_cur & ~31) | (src & 31)]; w_sync(); return r; }
int w_shfl(int v, int src) { g_xi[g_cur] = v; w_sync(); const int r = g_xi[(g_cur & ~31) | (src & 31)]; w_sync(); return r; }
double w_shfl_xor(double v, int mask) { return w_shfl(v, (g_cur & 31) ^ mask); }
int w_shfl_xor(int v, int mask) { return w_shfl(v, (g_cur & 31) ^ mask); }
unsigned w_ballot(int pred)
{
  g_xi[g_cur] = pred ? 1 : 0;
  w_sync();
  unsigned m = 0u;
  const int base = g_cur & ~31;
  for (int i = 0; i < 32; ++i) if (g_xi[base + i]) m |= 1u << i;
  w_sync();
  return m;
}

#include "kernel/rollout.cuh"
#include "kernel/sweep.cuh"

static void fiberEntry()
{
  (*g_body)();
  g_done[g_cur] = true;
  swapcontext(&g_ctx[g_cur], &g_main);
}

static void runBlock(int nWarps, std::function<void()> body);
static void runWarp(std::function<void()> body) { runBlock(1, body); }

static void runBlock(int nWarps, std::function<void()> body)
{
  const size_t stackSize = 512 * 1024;
  const int nFib = 32 * nWarps;
  if (g_stacks.size() < nFib * stackSize) g_stacks.resize(nFib * stackSize);
  g_body = &body;
  for (int l = 0; l < nFib; ++l) {
    g_done[l] = false;
    getcontext(&g_ctx[l]);
    g_ctx[l].uc_stack.ss_sp = g_stacks.data() + l * stackSize;
    g_ctx[l].uc_stack.ss_size = stackSize;
    g_ctx[l].uc_link = &g_main;
    makecontext(&g_ctx[l], fiberEntry, 0);
  }
  bool any = true;
  while (any) {
    any = false;
    for (int l = 0; l < nFib; ++l) {
      if (g_done[l]) continue;
      g_cur = l;
      swapcontext(&g_main, &g_ctx[l]);
      any = true;
    }
  }
}

extern "C" void aff_flop_count_rollout(const void* plan, int cand, unsigned long long out[5]);   // flop_count.cpp
static unsigned long long* g_flops = nullptr;   // when set, emu_run(2, ...) counts instead of predicting

static int emu_run(int sweep, const aff_scene_desc* scene, const aff_task_desc* tasks, size_t nTasks,
                                 const aff_constraint* cons, size_t nCons, const aff_candidate* cands, size_t nCands,
                                 const aff_opts* opts, aff_result* out, double* qlog, size_t qlogRows, char* errbuf, size_t errlen)
{
  affk::SceneModel model;
  std::string e = model.init(scene);
  affk::HostPlan H;
  int rc = e.empty() ? AFF_OK : AFF_ERR_INVALID;
  if (rc == AFF_OK) rc = affk::buildPlan(model, tasks, nTasks, cons, nCons, cands, nCands, *opts, H, e);
  if (rc != AFF_OK) { if (errbuf && errlen) { std::strncpy(errbuf, e.c_str(), errlen - 1); errbuf[errlen - 1] = 0; } return rc; }
  KPlan P = H.plan;
  std::vector<double> statA(12 * (size_t)(P.nStatic + 1)), statSC(2 * (size_t)P.nStatic + 2), dynSC(2 * (size_t)P.nDyn + 2),
      sshapeA(12 * (size_t)P.nStatShapes + 12), sbodyAABB(6 * (size_t)P.nStatBodies + 6), sseg(8 * (size_t)P.nStatShapes + 8);
  std::vector<aff_result> results(nCands);
  P.jq_init = H.jq_init.data(); P.jq0 = H.jq0.data(); P.jq_min = H.jq_min.data(); P.jq_max = H.jq_max.data();
  P.jspeed = H.jspeed.data(); P.jacc = H.jacc.data(); P.jwjl = H.jwjl.data(); P.jwmetric = H.jwmetric.data();
  P.j_node = H.j_node.data(); P.j_type = H.j_type.data(); P.dyn = H.dyn.data(); P.stat = H.stat.data();
  P.statA = statA.data(); P.statSC = statSC.data(); P.dynSC = dynSC.data(); P.fktab = H.fktab.data(); P.statpar_ref = H.statpar_ref.data(); P.obj_parent_slot = H.obj_parent_slot.data(); P.obj_parent_tree = H.obj_parent_tree.data(); P.dyn_under = H.dyn_under.data(); P.dyn_chain = H.dyn_chain.data(); P.dynT = H.dynT.data(); P.dynQ = H.dynQ.data();
  P.dshape = H.dshape.data(); P.dbody = H.dbody.data(); P.sshape = H.sshape.data(); P.sbody = H.sbody.data();
  P.sshapeA = sshapeA.data(); P.sbodyAABB = sbodyAABB.data(); P.pairsSS = H.pairsSS.data(); P.pairsSB = H.pairsSB.data(); P.sseg = sseg.data();
  P.obj_node = H.obj_node.data(); P.obj_body = H.obj_body.data(); P.obj_q6 = H.obj_q6.data(); P.dyn_pslot = H.dyn_pslot.data(); P.sweep_order = H.sweep_order.data(); P.node_shape_off = H.node_shape_off.data(); P.node_shape_idx = H.node_shape_idx.data(); P.dyn_store = H.dyn_store.data();
  P.tasks = H.tasks.data(); P.vias = H.vias.data(); P.acts = H.acts.data(); P.gcons = H.gcons.data(); P.cands = H.cands.data();
  P.results = results.data();
  P.qlog = qlog; P.qlog_rows = qlog ? (int)qlogRows : 0;
  runWarp([&]() { k_prologue(P); });
  if (sweep == 2) {
    // FP64 operation counts of every candidate (add, mul, fma, div, sqrt), kernel source with a counting scalar
    if (!P.sweep_ok) { if (errbuf && errlen) std::strncpy(errbuf, "batch exceeds the sweep kernel's caps", errlen - 1); return AFF_ERR_CAPACITY; }
    for (size_t i = 0; i < nCands; ++i) aff_flop_count_rollout(&P, (int)i, g_flops + 5 * i);
    for (size_t i = 0; i < nCands; ++i) out[i] = results[i];
    return AFF_OK;
  }
  if (sweep) {
    // thread-per-candidate kernel: plain sequential code, one call per candidate in the plan's order
    if (!P.sweep_ok) { if (errbuf && errlen) std::strncpy(errbuf, "batch exceeds the sweep kernel's caps", errlen - 1); return AFF_ERR_CAPACITY; }
    std::vector<TState> S(1);
    for (size_t i = 0; i < nCands; ++i) {
      std::memset(&S[0], 0, sizeof(TState));
      t_rollout(P, P.sweep_order[i], S[0], true);
    }
    for (size_t i = 0; i < nCands; ++i) out[i] = results[i];
    return AFF_OK;
  }
  if (sweep == 3) {
    // block-per-candidate kernel: warp 0 = step loop, warps 1 .. AFFC_WORKERS = collision model, last warp = trajectories
    std::vector<double> sm((size_t)P.warp_doubles * AFFC_WARPS + 64);
    for (size_t i = 0; i < nCands; ++i) {
      std::fill(sm.begin(), sm.end(), 0.0);
      KCtaShared X;
      std::memset(&X, 0, sizeof(X));
      runBlock(AFFC_WARPS, [&]() {
        const int w = emu_warp();
        double* slab = sm.data() + (size_t)w * P.warp_doubles;
        if (w == 0) k_rollout_t<1>(P, (int)i, slab, &X);
        else if (w <= AFFC_WORKERS) k_cta_worker(P, (int)i, slab, sm.data() + P.o_node, &X, w);
        else if (w == AFFC_WORKERS + 1) k_cta_traj(P, (int)i, slab, sm.data() + k_xcur_offset(P), &X);
        else if (w == AFFC_FK_WARP) k_cta_fk(P, (int)i, sm.data(), &X);
        else if (w == AFFC_GEO_WARP) k_cta_geometry(P, (int)i, sm.data(), &X);
        else if (w == AFFC_NS_WARP) k_cta_nullspace(P, (int)i, sm.data(), &X);
      });
      out[i] = results[i];
    }
    return AFF_OK;
  }
  std::vector<double> sm((size_t)P.warp_doubles + 64);
  for (size_t i = 0; i < nCands; ++i) {
    std::fill(sm.begin(), sm.end(), 0.0);
    runWarp([&]() { k_rollout(P, (int)i, sm.data()); });
    out[i] = results[i];
  }
  if (getenv("AFF_EMU_VERBOSE"))
    std::fprintf(stderr, "emu: nDyn %d nStatic %d rounds %d (%d for the step loop) dynShapes %d dynBodies %d statBodies %d pairs SS %d SB %d warp smem %zu B\n",
                 P.nDyn, P.nStatic, P.nRounds, P.nRoundsCrit, P.nDynShapes, P.nDynBodies, P.nStatBodies, P.nSS, P.nSB, H.smemBytesPerWarp);
  return AFF_OK;
}

extern "C" int emu_predict_batch(const aff_scene_desc* scene, const aff_task_desc* tasks, size_t nTasks,
                                 const aff_constraint* cons, size_t nCons, const aff_candidate* cands, size_t nCands,
                                 const aff_opts* opts, aff_result* out, double* qlog, size_t qlogRows, char* errbuf, size_t errlen)
{
  return emu_run(0, scene, tasks, nTasks, cons, nCons, cands, nCands, opts, out, qlog, qlogRows, errbuf, errlen);
}

// plan compiler only: {nDyn, nRounds, nRoundsCrit, warp_ok, sweep_ok, nStatLevels}; and, per dynamic node, its FK round
extern "C" int emu_plan_info(const aff_scene_desc* scene, const aff_task_desc* tasks, size_t nTasks,
                             const aff_constraint* cons, size_t nCons, const aff_candidate* cands, size_t nCands,
                             const aff_opts* opts, int* info, int* roundOf, size_t nRoundOf, int* parentOf)
{
  affk::SceneModel model;
  std::string e = model.init(scene);
  affk::HostPlan H;
  if (!e.empty()) return AFF_ERR_INVALID;
  const int rc = affk::buildPlan(model, tasks, nTasks, cons, nCons, cands, nCands, *opts, H, e);
  if (rc != AFF_OK) return rc;
  info[0] = H.plan.nDyn; info[1] = H.plan.nRounds; info[2] = H.plan.nRoundsCrit; info[3] = H.plan.warp_ok; info[4] = H.plan.sweep_ok; info[5] = H.plan.nStatLevels;
  for (size_t n = 0; n < H.roundOf.size() && n < nRoundOf; ++n) { roundOf[n] = H.roundOf[n]; parentOf[n] = H.dyn[n].parent; }
  return AFF_OK;
}

// the block-per-candidate arrangement (warp 0 + collision workers) on the CPU
extern "C" int emu_predict_batch_cta(const aff_scene_desc* scene, const aff_task_desc* tasks, size_t nTasks,
                                     const aff_constraint* cons, size_t nCons, const aff_candidate* cands, size_t nCands,
                                     const aff_opts* opts, aff_result* out, double* qlog, size_t qlogRows, char* errbuf, size_t errlen)
{
  return emu_run(3, scene, tasks, nTasks, cons, nCons, cands, nCands, opts, out, qlog, qlogRows, errbuf, errlen);
}

// the thread-per-candidate kernel source (sweep.cuh) on the CPU
extern "C" int emu_predict_batch_sweep(const aff_scene_desc* scene, const aff_task_desc* tasks, size_t nTasks,
                                       const aff_constraint* cons, size_t nCons, const aff_candidate* cands, size_t nCands,
                                       const aff_opts* opts, aff_result* out, double* qlog, size_t qlogRows, char* errbuf, size_t errlen)
{
  return emu_run(1, scene, tasks, nTasks, cons, nCons, cands, nCands, opts, out, qlog, qlogRows, errbuf, errlen);
}

// FP64 operation counts per candidate: counts[5 * i + {0: add/sub, 1: mul, 2: fma, 3: div, 4: sqrt}]
extern "C" int emu_count_flops(const aff_scene_desc* scene, const aff_task_desc* tasks, size_t nTasks,
                               const aff_constraint* cons, size_t nCons, const aff_candidate* cands, size_t nCands,
                               const aff_opts* opts, aff_result* out, unsigned long long* counts, char* errbuf, size_t errlen)
{
  g_flops = counts;
  const int rc = emu_run(2, scene, tasks, nTasks, cons, nCons, cands, nCands, opts, out, nullptr, 0, errbuf, errlen);
  g_flops = nullptr;
  return rc;
}
