// C ABI of the B200 predictor (include/affpredict.h): device memory, uploads,
// kernel launches.  There is no CPU fallback: every entry point that computes
// returns AFF_ERR_CUDA if the CUDA runtime / a B200 is not usable.
#include <cuda_runtime.h>

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdlib>
#include <cfloat>
#include <cstdio>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "affpredict.h"
#include "kernel/plan_build.h"
#include "aff_detmath.h"
#include "kernel/plan.h"
#include "kernel/warp.cuh"
// The kernel source twice (see the head of rollout.cuh): offsets read from the plan, and the fixed layout
// with compile-time offsets that the plan compiler picks for every batch within its caps.
#ifndef AFF_SWEEP_NO_BLOCK_SYNC
#define AFFT_BLOCK_SYNC 1
#endif
namespace kgeneric {
#include "kernel/rollout.cuh"
#include "kernel/sweep.cuh"
}
#define AFFK_FIXED_LAYOUT 1
namespace kfixed {
#include "kernel/rollout.cuh"
}
#undef AFFK_FIXED_LAYOUT

#ifndef AFF_SWEEP_MIN_DEFAULT
#define AFF_SWEEP_MIN_DEFAULT 24576
#endif
#ifndef AFF_WARPS_PER_BLOCK
#define AFF_WARPS_PER_BLOCK 4
#endif
#ifndef AFF_BLOCKS_PER_SM
#define AFF_BLOCKS_PER_SM 4     // register budget: 65536 / (4 * 128) = 128 per thread (5 blocks at 96 registers spill too much: measured slower)
#endif

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) { g_err = msg; return code; }
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) return fail(AFF_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_)); } while (0)

// ------------------------------------------------------------------ kernels
// Once per batch: world transforms of the static nodes, world frames / segments / AABBs of the static
// shapes, sin / cos of the constant joints.  One block; the static nodes go a tree level at a time, two nodes per
// warp and pass (12 lanes per node, as the FK of the step loop), the rest one item per thread.  Same
// operations per node as k_prologue (rollout.cuh: one warp, node after node; the CPU emulator runs that one).
#define AFF_PROLOGUE_THREADS 512
__global__ void __launch_bounds__(AFF_PROLOGUE_THREADS) aff_prologue_kernel(const KPlan P)
{
  const int tid = threadIdx.x, nT = blockDim.x, lane = tid & 31, warp = tid >> 5, nW = nT >> 5;
  for (int n = tid; n < P.nStatic; n += nT) {
    double s = 0.0, co = 1.0;
    if (P.stat[n].kind == AFFK_NODE_JOINT_CONST && P.stat[n].jtype >= AFF_JNT_ROT_X) aff_sincos(P.stat[n].qconst, &s, &co);
    P.statSC[2 * n] = s; P.statSC[2 * n + 1] = co;
  }
  for (int n = tid; n < P.nDyn; n += nT) {
    double s = 0.0, co = 1.0;
    if (P.dyn[n].kind == AFFK_NODE_JOINT_CONST && P.dyn[n].jtype >= AFF_JNT_ROT_X) aff_sincos(P.dyn[n].qconst, &s, &co);
    P.dynSC[2 * n] = s; P.dynSC[2 * n + 1] = co;
  }
  if (tid < 12) P.statA[12 * P.nStatic + tid] = (tid == 3 || tid == 7 || tid == 11) ? 1.0 : 0.0;
  __syncthreads();
  const int grp = lane / 12, e = lane % 12;      // lanes 24 .. 31 idle (they take part in the shuffle only)
  for (int l = 0; l < P.nStatLevels; ++l) {
    const int lo = P.stat_level_off[l], hi = P.stat_level_off[l + 1];
    for (int base = lo + 2 * warp; base < hi; base += 2 * nW) {
      const int idx = base + grp;
      const bool on = grp < 2 && idx < hi;
      const int n = P.stat_order[on ? idx : lo];
      const KNode* nd = P.stat + n;
      const int par = nd->parent;
      const double* Pp = (par == AFFK_WORLD) ? P.statA + 12 * P.nStatic : P.statA + 12 * AFFK_STATIC_IDX(par);
      double val = nd->identityT ? Pp[e] : kgeneric::k_compose_elem(nd->T, Pp, e);
      int partner = e, isA = 0, isB = 0;
      const int kind = nd->kind, jt = nd->jtype;
      if (kind == AFFK_NODE_JOINT_CONST) {
        if (jt >= AFF_JNT_ROT_X && e >= 3) {
          const int d = jt - AFF_JNT_ROT_X, a = (d + 1) % 3, b = (d + 2) % 3, row = (e - 3) / 3, col = (e - 3) % 3;
          if (row == a) { isA = 1; partner = 3 + 3 * b + col; } else if (row == b) { isB = 1; partner = 3 + 3 * a + col; }
        } else if (jt < AFF_JNT_ROT_X && e < 3) partner = 3 + 3 * jt + e;
      }
      const double other = __shfl_sync(0xffffffffu, val, (grp < 2 ? 12 * grp : 0) + partner);
      if (kind == AFFK_NODE_JOINT_CONST) {
        const double s = P.statSC[2 * n], co = P.statSC[2 * n + 1];
        if (jt >= AFF_JNT_ROT_X) { if (isA) val = fma(s, other, co * val); else if (isB) val = fma(co, val, -(s * other)); }
        else if (e < 3) val = fma(nd->qconst, other, val);
      }
      if (on) P.statA[12 * n + e] = val;
    }
    __syncthreads();
  }
  for (int s = tid; s < P.nStatShapes; s += nT) {
    const KShape* sh = P.sshape + s;
    const int ref = sh->node;
    const double* An = (ref == AFFK_WORLD) ? P.statA + 12 * P.nStatic : P.statA + 12 * AFFK_STATIC_IDX(ref);
    double W[12];
    kgeneric::k_compose(W, sh->A_CB, An);
    for (int k = 0; k < 12; ++k) P.sshapeA[12 * s + k] = W[k];
    const bool ssl = sh->type == AFF_SHAPE_SSL;
    for (int k = 0; k < 3; ++k) { P.sseg[8 * s + k] = W[k]; P.sseg[8 * s + 3 + k] = ssl ? sh->ext[2] * W[9 + k] : 0.0; }
    P.sseg[8 * s + 6] = sh->ext[0]; P.sseg[8 * s + 7] = sh->bound;
  }
  __syncthreads();
  for (int b = tid; b < P.nStatBodies; b += nT) {
    const KShapeBody* sb = P.sbody + b;
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    for (int k = 0; k < sb->n_shapes; ++k) {
      const KShape* sh = P.sshape + sb->first_shape + k;
      if (!sh->active0) continue;
      double smn[3], smx[3];
      kgeneric::k_shape_aabb(sh->type, sh->ext, P.sshapeA + 12 * (sb->first_shape + k), smn, smx);
      for (int q = 0; q < 3; ++q) { if (smn[q] < mn[q]) mn[q] = smn[q]; if (smx[q] > mx[q]) mx[q] = smx[q]; }
    }
    for (int q = 0; q < 3; ++q) { P.sbodyAABB[6 * b + q] = mn[q]; P.sbodyAABB[6 * b + 3 + q] = mx[q]; }
  }
}

// Persistent rollout kernel: one warp per candidate, warps stride over the batch.
#ifndef AFF_LAUNCH_BOUNDS
#define AFF_LAUNCH_BOUNDS __launch_bounds__(32 * AFF_WARPS_PER_BLOCK, AFF_BLOCKS_PER_SM)
#endif
#define AFF_ROLLOUT_KERNEL(NAME, NS)                                                                                       \
  __global__ void AFF_LAUNCH_BOUNDS NAME(const KPlan P)                      \
  {                                                                                                                      \
    extern __shared__ double smem[];                                                                                     \
    const int warp = threadIdx.x >> 5;                                                                                   \
    /* The warp's base is kept as an opaque 32-bit shared-memory address (otherwise it is re-derived from    \
       threadIdx and the shared window at every use, 13+ instructions each time); the pointer made from it   \
       is known to the compiler as a shared-memory pointer, so accesses stay LDS / STS. */                   \
    unsigned smAddr = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)warp * (unsigned)P.warp_doubles * 8u;        \
    asm volatile("" : "+r"(smAddr));                                                                                    \
    double* sm = (double*)__cvta_shared_to_generic((size_t)smAddr);                                                      \
    const int warpsTotal = gridDim.x * AFF_WARPS_PER_BLOCK;                                                              \
    for (int cand = blockIdx.x * AFF_WARPS_PER_BLOCK + warp; cand < P.nCands; cand += warpsTotal) {                      \
      NS::k_rollout(P, cand, sm);                                                                                        \
      __syncwarp();                                                                                                      \
    }                                                                                                                    \
  }
AFF_ROLLOUT_KERNEL(aff_rollout_kernel, kfixed)            // batches in the fixed layout (KPlan::fixed_layout)
AFF_ROLLOUT_KERNEL(aff_rollout_kernel_generic, kgeneric)  // anything larger

// Block-per-candidate kernel for batches far below one wave (the call sites that check a handful of
// candidates, or one trajectory): warp 0 walks the step loop, AFFC_WORKERS warps evaluate the collision
// model of the finished steps on private copies of the frames, a step or two behind, and one warp runs
// the time-driven part (trajectories, activation points, blending) ahead of it (k_rollout_t<1>,
// k_cta_worker, k_cta_traj in rollout.cuh).  One warp per scheduler; the rollout is a dependent chain, so the time
// of a step is what counts here, not the work per SM.
#ifndef AFF_CTA_BLOCKS_PER_SM
#define AFF_CTA_BLOCKS_PER_SM 1     // 255 registers per thread: the step loop without spills
#endif
#define AFF_CTA_THREADS (32 * AFFC_WARPS)
#define AFF_CTA_KERNEL(NAME, NS)                                                                                         \
  __global__ void __launch_bounds__(AFF_CTA_THREADS, AFF_CTA_BLOCKS_PER_SM) NAME(const KPlan P)                          \
  {                                                                                                                      \
    extern __shared__ double smem[];                                                                                     \
    const int warp = threadIdx.x >> 5;                                                                                   \
    unsigned smAddr = (unsigned)__cvta_generic_to_shared(smem) + (unsigned)warp * (unsigned)P.warp_doubles * 8u;        \
    asm volatile("" : "+r"(smAddr));                                                                                    \
    double* sm = (double*)__cvta_shared_to_generic((size_t)smAddr);                                                      \
    NS::KCtaShared* X = (NS::KCtaShared*)(smem + (size_t)(AFFC_WORKERS + 2) * P.warp_doubles);                           \
    for (int cand = blockIdx.x; cand < P.nCands; cand += gridDim.x) {                                                    \
      if (threadIdx.x == 0) {                                                                                            \
        X->fkSeq = 0; X->fk2Seq = 0; X->stop = 0; X->trajSeq = 0; X->geoSeq = 0; X->consSeq = 0; X->nsSeq = 0; X->trkSeq = 0; X->geoHelpSeq = 0;              \
        for (int w = 0; w < AFFC_WORKERS; ++w) { X->copied[w] = 0; X->done[w] = 0; }                                     \
      }                                                                                                                  \
      __syncthreads();                                                                                                   \
      if (warp == 0) NS::template k_rollout_t<1>(P, cand, sm, X);                                                        \
      else if (warp <= AFFC_WORKERS) NS::k_cta_worker(P, cand, sm, smem + P.o_node, X, warp);                            \
      else if (warp == AFFC_WORKERS + 1) NS::k_cta_traj(P, cand, sm, smem + NS::k_xcur_offset(P), X);                    \
      else if (warp == AFFC_FK_WARP) NS::k_cta_fk(P, cand, smem, X);                                                     \
      else if (warp == AFFC_GEO_WARP) NS::k_cta_geometry(P, cand, smem, X);                                              \
      else if (warp == AFFC_NS_WARP) NS::k_cta_nullspace(P, cand, smem, X);                                              \
      __syncthreads();                                                                                                   \
    }                                                                                                                    \
  }
AFF_CTA_KERNEL(aff_rollout_cta_kernel, kfixed)
AFF_CTA_KERNEL(aff_rollout_cta_kernel_generic, kgeneric)

// Thread-per-candidate kernel (sweep.cuh): 32 rollouts per warp.  Persistent blocks: block b takes the
// chunks b, b + gridDim, ... of THREADS consecutive entries of sweep_order (candidates grouped by
// structure), and its warps walk the step loop in lock-step phases (T_SYNC in sweep.cuh).  THREADS x
// blocks per SM = 1024 rollouts per SM at 64 registers per thread (measured round 2: 32 warps per SM in
// barrier groups of 16 beat 8 / 16 / 24 warps and smaller groups); the smaller block sizes spread
// batches below one wave over all SMs.
template <int THREADS> __global__ void __launch_bounds__(THREADS, 1024 / THREADS) aff_sweep_kernel(const __grid_constant__ KPlan P)
{
  for (int base = blockIdx.x * THREADS; base < P.nCands; base += gridDim.x * THREADS) {
    const int i = base + threadIdx.x;
    // threads past the end of the batch shadow its last candidate (they take part in the block's
    // barriers and write nothing)
    const bool alive = i < P.nCands;
    kgeneric::TState S;
    kgeneric::t_rollout(P, __ldg(&P.sweep_order[alive ? i : P.nCands - 1]), S, alive);
    __syncthreads();
  }
}

// Register-resident DFMA chains: the FP64 roofline denominator.
__global__ void aff_dfma_peak_kernel(double* out, int iters)
{
  double a0 = threadIdx.x * 1e-9 + 1.0, a1 = a0 + 0.1, a2 = a0 + 0.2, a3 = a0 + 0.3, a4 = a0 + 0.4, a5 = a0 + 0.5, a6 = a0 + 0.6, a7 = a0 + 0.7;
  const double b = 0.999999, cc = 1e-7;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, cc); a1 = fma(a1, b, cc); a2 = fma(a2, b, cc); a3 = fma(a3, b, cc);
    a4 = fma(a4, b, cc); a5 = fma(a5, b, cc); a6 = fma(a6, b, cc); a7 = fma(a7, b, cc);
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

// ------------------------------------------------------------------ handles
struct aff_scene
{
  int device = 0;
  int smCount = 0, smemOptin = 0;     // queried once: cudaGetDeviceProperties costs milliseconds per call
  affk::SceneModel model;
};

template <class T> struct DevBuf
{
  T* p = nullptr; size_t n = 0;
  cudaError_t upload(const std::vector<T>& v, cudaStream_t s)
  {
    n = v.size();
    cudaError_t e = cudaMallocAsync((void**)&p, std::max<size_t>(1, n) * sizeof(T), s);
    if (e != cudaSuccess) return e;
    if (n) e = cudaMemcpyAsync(p, v.data(), n * sizeof(T), cudaMemcpyHostToDevice, s);
    return e;
  }
  cudaError_t alloc(size_t count)
  {
    n = count;
    return cudaMallocAsync((void**)&p, std::max<size_t>(1, n) * sizeof(T), 0);
  }
  // Stream-ordered allocations from the device's default memory pool (kept, not returned to the driver:
  // see aff_scene_create), so that a predict call per action does not pay cudaMalloc / cudaFree each time.
  void release() { if (p) cudaFreeAsync(p, 0); p = nullptr; }
};

struct aff_batch
{
  aff_scene* scene = nullptr;
  affk::HostPlan host;
  KPlan plan{};
  DevBuf<double> jq_init, jq0, jq_min, jq_max, jspeed, jacc, jwjl, jwmetric, obj_q6, statA, statSC, dynSC, sshapeA, sbodyAABB, qlog;
  DevBuf<KFkEntry> fktab;
  DevBuf<int32_t> j_node, j_type, statpar_ref, obj_parent_slot, obj_parent_tree, dyn_under, obj_node, obj_body, dyn_pslot, sweep_order, node_shape_off, node_shape_idx, dyn_store, stat_order, stat_level_off;
  DevBuf<uint32_t> dyn_chain;
  DevBuf<double> dynT, dynQ;
  DevBuf<KPair> pairsSS, pairsSB;
  DevBuf<double> sseg;
  DevBuf<KNode> dyn, stat;
  DevBuf<KShape> dshape, sshape;
  DevBuf<KShapeBody> dbody, sbody;
  DevBuf<KTask> tasks; DevBuf<KVia> vias; DevBuf<KAct> acts; DevBuf<KGraphCon> gcons; DevBuf<KCand> cands;
  DevBuf<aff_result> results;
  int grid = 0; size_t smemBytes = 0; int launches = 0;
  size_t h2dBytes = 0;
  bool prologueDone = false, useFixed = false, useSweep = false, useCta = false;
  int ctaGrid = 0; size_t ctaSmemBytes = 0;
  int sweepGrid = 0, sweepThreads = 512;
  cudaStream_t lastStream = nullptr; bool launched = false;
  void release()
  {
    jq_init.release(); jq0.release(); jq_min.release(); jq_max.release(); jspeed.release(); jacc.release(); jwjl.release();
    jwmetric.release(); obj_q6.release(); statA.release(); statSC.release(); dynSC.release(); sshapeA.release();
    sbodyAABB.release(); qlog.release(); j_node.release(); j_type.release(); fktab.release(); statpar_ref.release(); obj_parent_slot.release(); obj_parent_tree.release(); dyn_under.release(); dyn_chain.release(); dynT.release(); dynQ.release(); pairsSS.release(); pairsSB.release(); sseg.release();
    obj_node.release(); obj_body.release(); dyn_pslot.release(); sweep_order.release(); node_shape_off.release(); node_shape_idx.release(); dyn_store.release(); stat_order.release(); stat_level_off.release(); dyn.release(); stat.release(); dshape.release(); sshape.release();
    dbody.release(); sbody.release(); tasks.release(); vias.release(); acts.release(); gcons.release(); cands.release();
    results.release();
  }
};

extern "C" {

const char* aff_last_error(void) { return g_err.c_str(); }

int aff_device_count(void)
{
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) { g_err = "cudaGetDeviceCount failed: no usable CUDA device"; return AFF_ERR_CUDA; }
  return n;
}

int aff_scene_create(const aff_scene_desc* desc, int device, aff_scene** out)
{
  if (!desc || !out) return fail(AFF_ERR_INVALID, "aff_scene_create: null argument");
  int n = 0;
  CUDA_TRY(cudaGetDeviceCount(&n));
  if (device < 0 || device >= n) return fail(AFF_ERR_CUDA, "aff_scene_create: CUDA device " + std::to_string(device) + " not present (no CPU fallback)");
  aff_scene* s = new aff_scene;
  s->device = device;
  if (cudaDeviceGetAttribute(&s->smCount, cudaDevAttrMultiProcessorCount, device) != cudaSuccess ||
      cudaDeviceGetAttribute(&s->smemOptin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device) != cudaSuccess) {
    delete s;
    return fail(AFF_ERR_CUDA, "aff_scene_create: cannot query the device");
  }
  {
    cudaMemPool_t pool;
    unsigned long long keep = ~0ull;
    if (cudaSetDevice(device) == cudaSuccess && cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess)
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
  }
  const std::string e = s->model.init(desc);
  if (!e.empty()) { delete s; return fail(AFF_ERR_INVALID, "aff_scene_create: " + e); }
  *out = s;
  return AFF_OK;
}

void aff_scene_destroy(aff_scene* s) { delete s; }
int aff_scene_num_ik_joints(const aff_scene* s) { return s ? s->model.nJ : 0; }

static int uploadBatch(aff_batch* b, cudaStream_t st)
{
  affk::HostPlan& H = b->host;
  KPlan& P = b->plan;
  P = H.plan;
#define UP(name) do { CUDA_TRY(b->name.upload(H.name, st)); b->h2dBytes += H.name.size() * sizeof(H.name[0]); } while (0)
  UP(jq_init); UP(jq0); UP(jq_min); UP(jq_max); UP(jspeed); UP(jacc); UP(jwjl); UP(jwmetric); UP(obj_q6);
  UP(j_node); UP(j_type); UP(fktab); UP(statpar_ref); UP(obj_parent_slot); UP(obj_parent_tree); UP(dyn_under); UP(dyn_chain); UP(dynT); UP(dynQ); UP(pairsSS); UP(pairsSB); UP(obj_node); UP(obj_body); UP(dyn_pslot); UP(sweep_order); UP(node_shape_off); UP(node_shape_idx); UP(dyn_store); UP(stat_order); UP(stat_level_off);
  UP(dyn); UP(stat); UP(dshape); UP(sshape); UP(dbody); UP(sbody);
  UP(tasks); UP(vias); UP(acts); UP(gcons); UP(cands);
#undef UP
  CUDA_TRY(b->statA.alloc(12 * (size_t)(P.nStatic + 1)));
  CUDA_TRY(b->statSC.alloc(2 * (size_t)std::max(P.nStatic, 1)));
  CUDA_TRY(b->dynSC.alloc(2 * (size_t)std::max(P.nDyn, 1)));
  CUDA_TRY(b->sshapeA.alloc(12 * (size_t)std::max(P.nStatShapes, 1)));
  CUDA_TRY(b->sbodyAABB.alloc(6 * (size_t)std::max(P.nStatBodies, 1)));
  CUDA_TRY(b->sseg.alloc(8 * (size_t)std::max(P.nStatShapes, 1)));
  CUDA_TRY(b->results.alloc((size_t)P.nCands));
  P.qlog = nullptr; P.qlog_rows = 0;
  if (P.opts.log_q) {
    P.qlog_rows = H.maxSteps + 1;
    CUDA_TRY(b->qlog.alloc((size_t)P.nCands * P.qlog_rows * P.nJ));
    // rows a rollout never reaches (initial state invalid) read as zero
    CUDA_TRY(cudaMemsetAsync(b->qlog.p, 0, sizeof(double) * (size_t)P.nCands * P.qlog_rows * P.nJ, st));
    P.qlog = b->qlog.p;
  }
  P.jq_init = b->jq_init.p; P.jq0 = b->jq0.p; P.jq_min = b->jq_min.p; P.jq_max = b->jq_max.p; P.jspeed = b->jspeed.p;
  P.jacc = b->jacc.p; P.jwjl = b->jwjl.p; P.jwmetric = b->jwmetric.p; P.j_node = b->j_node.p; P.j_type = b->j_type.p;
  P.dyn = b->dyn.p; P.stat = b->stat.p; P.statA = b->statA.p; P.statSC = b->statSC.p; P.dynSC = b->dynSC.p; P.fktab = b->fktab.p; P.statpar_ref = b->statpar_ref.p; P.obj_parent_slot = b->obj_parent_slot.p; P.obj_parent_tree = b->obj_parent_tree.p; P.dyn_under = b->dyn_under.p; P.dyn_chain = b->dyn_chain.p; P.dynT = b->dynT.p; P.dynQ = b->dynQ.p;
  P.dshape = b->dshape.p; P.dbody = b->dbody.p; P.sshape = b->sshape.p; P.sbody = b->sbody.p; P.sshapeA = b->sshapeA.p;
  P.sbodyAABB = b->sbodyAABB.p; P.pairsSS = b->pairsSS.p; P.pairsSB = b->pairsSB.p; P.sseg = b->sseg.p; P.obj_node = b->obj_node.p; P.obj_body = b->obj_body.p; P.obj_q6 = b->obj_q6.p; P.dyn_pslot = b->dyn_pslot.p; P.sweep_order = b->sweep_order.p; P.node_shape_off = b->node_shape_off.p; P.node_shape_idx = b->node_shape_idx.p; P.dyn_store = b->dyn_store.p; P.stat_order = b->stat_order.p; P.stat_level_off = b->stat_level_off.p;
  P.tasks = b->tasks.p; P.vias = b->vias.p; P.acts = b->acts.p; P.gcons = b->gcons.p; P.cands = b->cands.p;
  P.results = b->results.p;
  return AFF_OK;
}

int aff_batch_create(aff_scene* s, const aff_task_desc* tasks, size_t n_tasks, const aff_constraint* cons, size_t n_cons,
                     const aff_candidate* cands, size_t n_cands, const aff_opts* opts, aff_batch** out)
{
  if (!s || !out || !cands || (!tasks && n_tasks) || (!cons && n_cons)) return fail(AFF_ERR_INVALID, "aff_batch_create: null argument");
  aff_opts o;
  if (opts) o = *opts; else aff_default_opts(&o);
  if (!(o.dt > 0.0)) return fail(AFF_ERR_INVALID, "aff_batch_create: dt must be positive");
  CUDA_TRY(cudaSetDevice(s->device));
  // owned until the end: every early return below frees the device buffers and the handle
  struct Drop { void operator()(aff_batch* p) const { p->release(); delete p; } };
  std::unique_ptr<aff_batch, Drop> owner(new aff_batch);
  aff_batch* b = owner.get();
  b->scene = s;
  std::string e;
  const auto tb0 = std::chrono::steady_clock::now();
  int rc = affk::buildPlan(s->model, tasks, n_tasks, cons, n_cons, cands, n_cands, o, b->host, e);
  if (rc != AFF_OK) return fail(rc, "aff_batch_create: " + e);
  const auto tb1 = std::chrono::steady_clock::now();
  rc = uploadBatch(b, 0);
  if (getenv("AFF_TIMING")) {
    cudaStreamSynchronize(0);
    const auto tb2 = std::chrono::steady_clock::now();
    std::fprintf(stderr, "aff_batch_create: plan %.1f ms, alloc + H2D %.1f ms (%.1f MB)\n", std::chrono::duration<double, std::milli>(tb1 - tb0).count(),
                 std::chrono::duration<double, std::milli>(tb2 - tb1).count(), b->h2dBytes / 1e6);
  }
  if (rc != AFF_OK) return rc;
  // Kernel choice: one rollout per thread for batches that fit its per-thread caps and have at least
  // AFF_SWEEP_MIN candidates (default below: where it overtakes one rollout per warp, which finishes a
  // wave of 2368 in 67 ms but needs 9.4 k warp instructions per rollout step); AFF_KERNEL=warp / sweep
  // forces one of the two (the parity tests run both).
  {
    const char* force = getenv("AFF_KERNEL");
    const char* mn = getenv("AFF_SWEEP_MIN");
    const size_t minSweep = mn ? (size_t)atoll(mn) : AFF_SWEEP_MIN_DEFAULT;
    b->useSweep = b->plan.sweep_ok && n_cands >= minSweep;
    if (force && !strcmp(force, "warp")) b->useSweep = false;
    if (force && !strcmp(force, "sweep")) {
      if (!b->plan.sweep_ok) return fail(AFF_ERR_CAPACITY, "aff_batch_create: AFF_KERNEL=sweep but the batch exceeds the per-thread caps");
      b->useSweep = true;
    }
    // block size: the largest that still gives every SM a block; grid: at most one resident wave
    b->sweepThreads = 512;
    while (b->sweepThreads > 128 && (n_cands + b->sweepThreads - 1) / b->sweepThreads < (size_t)s->smCount) b->sweepThreads /= 2;
    if (const char* th = getenv("AFF_SWEEP_THREADS")) { const int v = atoi(th); if (v == 128 || v == 256 || v == 512) b->sweepThreads = v; }
    const size_t blocksNeededS = (n_cands + b->sweepThreads - 1) / b->sweepThreads;
    if (!b->plan.warp_ok && !b->useSweep) {
      if (!b->plan.sweep_ok || (force && !strcmp(force, "warp")))
        return fail(AFF_ERR_UNSUPPORTED, "aff_batch_create: an object is connected to a body the warp-per-candidate FK schedule computes after it, "
                                         "and the batch does not fit the thread-per-candidate kernel");
      b->useSweep = true;
    }
    b->sweepGrid = (int)std::min(blocksNeededS, (size_t)s->smCount * (size_t)(1024 / b->sweepThreads));
    if (b->useSweep) {
      // the kernel uses no shared memory to speak of: give the whole unified array to L1 (the per-thread state lives in local memory)
      const char* co = getenv("AFF_TUNE_SWEEP_CARVEOUT");
      const int carve = co ? atoi(co) : 0;
      const void* kf = b->sweepThreads == 512 ? (const void*)aff_sweep_kernel<512> : (b->sweepThreads == 256 ? (const void*)aff_sweep_kernel<256> : (const void*)aff_sweep_kernel<128>);
      CUDA_TRY(cudaFuncSetAttribute(kf, cudaFuncAttributePreferredSharedMemoryCarveout, carve));
    }
  }
  // launch geometry: persistent blocks, as many as fit per SM
  b->smemBytes = b->host.smemBytesPerWarp * AFF_WARPS_PER_BLOCK;
  // tuning experiments only (tools/try_variants.sh): extra dynamic shared memory per block, preferred carve-out in percent
  if (const char* pad = getenv("AFF_TUNE_SMEM_PAD")) b->smemBytes += (size_t)atoi(pad);
  if (b->smemBytes > (size_t)s->smemOptin) return fail(AFF_ERR_CAPACITY, "aff_batch_create: rollout state does not fit in shared memory");
  // AFF_TUNE_GENERIC_KERNEL: run the generic kernel on a fixed-layout plan too (A/B of the two builds)
  b->useFixed = b->plan.fixed_layout != 0 && !getenv("AFF_TUNE_GENERIC_KERNEL");
  const void* kfn = b->useFixed ? (const void*)aff_rollout_kernel : (const void*)aff_rollout_kernel_generic;
  {
    // The limit is a property of the kernel, not of the batch: only ever raise it, or a batch created
    // earlier with a larger footprint could no longer be launched.
    static std::mutex mtx;
    static size_t raised[64][2] = {};
    std::lock_guard<std::mutex> lock(mtx);
    size_t& cur = raised[s->device & 63][b->useFixed ? 0 : 1];
    if (b->smemBytes > cur) {
      CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->smemBytes));
      cur = b->smemBytes;
    }
  }
  if (const char* co = getenv("AFF_TUNE_CARVEOUT")) CUDA_TRY(cudaFuncSetAttribute(kfn, cudaFuncAttributePreferredSharedMemoryCarveout, atoi(co)));
  int perSM = 0;
  CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSM, kfn, 32 * AFF_WARPS_PER_BLOCK, b->smemBytes));
  if (perSM < 1) perSM = 1;
  const int blocksNeeded = (int)((n_cands + AFF_WARPS_PER_BLOCK - 1) / AFF_WARPS_PER_BLOCK);
  b->grid = std::max(1, std::min(blocksNeeded, perSM * s->smCount));
  // block-per-candidate kernel: batches of at most AFF_CTA_MAX candidates (default: what one wave of
  // its blocks holds); AFF_KERNEL=cta forces it, AFF_KERNEL=warp / sweep rule it out
  {
    const char* force = getenv("AFF_KERNEL");
    b->ctaSmemBytes = b->host.smemBytesPerWarp * (AFFC_WORKERS + 2) + sizeof(kgeneric::KCtaShared);     // slabs: step loop, workers, trajectories
    const void* cfn = b->useFixed ? (const void*)aff_rollout_cta_kernel : (const void*)aff_rollout_cta_kernel_generic;
    const bool fits = b->plan.warp_ok && b->ctaSmemBytes <= (size_t)s->smemOptin;
    int perSMc = 0;
    if (fits) {
      static std::mutex mtx;
      static size_t raised[64][2] = {};
      std::lock_guard<std::mutex> lock(mtx);
      size_t& cur = raised[s->device & 63][b->useFixed ? 0 : 1];
      if (b->ctaSmemBytes > cur) {
        CUDA_TRY(cudaFuncSetAttribute(cfn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)b->ctaSmemBytes));
        cur = b->ctaSmemBytes;
      }
      CUDA_TRY(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSMc, cfn, AFF_CTA_THREADS, b->ctaSmemBytes));
    }
    const char* mx = getenv("AFF_CTA_MAX");
    const size_t ctaMax = mx ? (size_t)atoll(mx) : (size_t)std::max(perSMc, 0) * (size_t)s->smCount;
    b->useCta = fits && perSMc >= 1 && !b->useSweep && n_cands <= ctaMax;
    if (force && !strcmp(force, "warp")) b->useCta = false;
    if (force && !strcmp(force, "cta")) {
      if (!fits || perSMc < 1) return fail(AFF_ERR_CAPACITY, "aff_batch_create: AFF_KERNEL=cta but the batch does not fit the block-per-candidate kernel");
      b->useCta = true; b->useSweep = false;
    }
    b->ctaGrid = (int)std::max<size_t>(1, std::min(n_cands, (size_t)std::max(perSMc, 1) * (size_t)s->smCount));
  }
  CUDA_TRY(cudaStreamSynchronize(0));
  *out = owner.release();
  return AFF_OK;
}

void aff_batch_destroy(aff_batch* b)
{
  if (!b) return;
  // the buffers go back to the pool in stream order of the default stream: wait for the stream that used them
  if (b->launched) cudaStreamSynchronize(b->lastStream);
  b->release();
  delete b;
}

int aff_batch_launch(aff_batch* b, void* stream)
{
  if (!b) return fail(AFF_ERR_INVALID, "aff_batch_launch: null batch");
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaSetDevice(b->scene->device));
  b->launches = 0;
  b->lastStream = st; b->launched = true;
  if (!b->prologueDone) {
    aff_prologue_kernel<<<1, AFF_PROLOGUE_THREADS, 0, st>>>(b->plan);
    CUDA_TRY(cudaGetLastError());
    b->prologueDone = true;
    b->launches++;
  }
  if (b->useCta) {
    if (b->useFixed) aff_rollout_cta_kernel<<<b->ctaGrid, AFF_CTA_THREADS, b->ctaSmemBytes, st>>>(b->plan);
    else aff_rollout_cta_kernel_generic<<<b->ctaGrid, AFF_CTA_THREADS, b->ctaSmemBytes, st>>>(b->plan);
  }
  else if (b->useSweep) {
    if (b->sweepThreads == 512) aff_sweep_kernel<512><<<b->sweepGrid, 512, 0, st>>>(b->plan);
    else if (b->sweepThreads == 256) aff_sweep_kernel<256><<<b->sweepGrid, 256, 0, st>>>(b->plan);
    else aff_sweep_kernel<128><<<b->sweepGrid, 128, 0, st>>>(b->plan);
  }
  else if (b->useFixed) aff_rollout_kernel<<<b->grid, 32 * AFF_WARPS_PER_BLOCK, b->smemBytes, st>>>(b->plan);
  else aff_rollout_kernel_generic<<<b->grid, 32 * AFF_WARPS_PER_BLOCK, b->smemBytes, st>>>(b->plan);
  CUDA_TRY(cudaGetLastError());
  b->launches++;
  return AFF_OK;
}

void* aff_batch_results_device(aff_batch* b) { return b ? (void*)b->plan.results : nullptr; }

int aff_batch_set_results_buffer(aff_batch* b, void* device_ptr)
{
  if (!b) return fail(AFF_ERR_INVALID, "aff_batch_set_results_buffer: null batch");
  b->plan.results = device_ptr ? (aff_result*)device_ptr : b->results.p;
  return AFF_OK;
}

int aff_batch_results(aff_batch* b, void* stream, aff_result* out, size_t n)
{
  if (!b || !out) return fail(AFF_ERR_INVALID, "aff_batch_results: null argument");
  if (n > (size_t)b->plan.nCands) n = (size_t)b->plan.nCands;
  cudaStream_t st = (cudaStream_t)stream;
  CUDA_TRY(cudaMemcpyAsync(out, b->plan.results, n * sizeof(aff_result), cudaMemcpyDeviceToHost, st));
  CUDA_TRY(cudaStreamSynchronize(st));
  return AFF_OK;
}

int aff_batch_last_launch_count(const aff_batch* b) { return b ? b->launches : 0; }
int aff_batch_num_steps(const aff_batch* b, size_t cand) { return (b && cand < b->host.cands.size()) ? b->host.cands[cand].n_steps : -1; }
size_t aff_batch_h2d_bytes(const aff_batch* b) { return b ? b->h2dBytes : 0; }
size_t aff_batch_smem_bytes_per_warp(const aff_batch* b) { return b ? b->host.smemBytesPerWarp : 0; }
int aff_batch_grid_blocks(const aff_batch* b) { return b ? (b->useCta ? b->ctaGrid : (b->useSweep ? b->sweepGrid : b->grid)) : 0; }
int aff_batch_kernel_kind(const aff_batch* b) { return !b ? 0 : (b->useCta ? 2 : (b->useSweep ? 1 : 0)); }

int aff_batch_fetch_q(aff_batch* b, size_t cand, double* out, size_t max_rows)
{
  if (!b || !out) return fail(AFF_ERR_INVALID, "aff_batch_fetch_q: null argument");
  if (!b->plan.qlog) return fail(AFF_ERR_INVALID, "aff_batch_fetch_q: batch was created without opts.log_q");
  if (cand >= (size_t)b->plan.nCands) return fail(AFF_ERR_INVALID, "aff_batch_fetch_q: candidate out of range");
  const size_t rows = std::min<size_t>(max_rows, (size_t)b->host.cands[cand].n_steps + 1);
  CUDA_TRY(cudaMemcpy(out, b->qlog.p + cand * (size_t)b->plan.qlog_rows * b->plan.nJ, rows * b->plan.nJ * sizeof(double), cudaMemcpyDeviceToHost));
  return (int)rows;
}

int aff_predict_batch(aff_scene* s, const aff_task_desc* tasks, size_t n_tasks, const aff_constraint* cons, size_t n_cons,
                      const aff_candidate* cands, size_t n_cands, const aff_opts* opts, aff_result* out)
{
  static const bool timing = getenv("AFF_TIMING") != nullptr;   // AFF_TIMING=1: per-stage wall clock on stderr
  auto now = [] { return std::chrono::steady_clock::now(); };
  auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  if (!s || !out) return fail(AFF_ERR_INVALID, "aff_predict_batch: null argument");
  // One device batch tracks at most AFFK_MAX_OBJ re-parentable objects (they are explicit state of the
  // kernels).  A call whose candidates name more objects in total - e.g. the candidates of three different
  // `get` commands - is split into groups of candidates whose objects fit, each group is predicted as its
  // own batch, and the records come back in the caller's order.
  {
    auto objectsOf = [&](const aff_candidate& c, int32_t* objs) {
      int n = 0;
      auto add = [&](int32_t b) { for (int k = 0; k < n; ++k) if (objs[k] == b) return; if (n < 8) objs[n] = b; ++n; };
      if (c.pose_body >= 0) add(c.pose_body);
      if (c.first_constraint >= 0 && (size_t)(c.first_constraint + c.n_constraints) <= n_cons)
        for (int k = 0; k < c.n_constraints; ++k) {
          const aff_constraint& cc = cons[c.first_constraint + k];
          if (cc.kind == AFF_CON_CONNECT || cc.kind == AFF_CON_COLLISION) add(cc.body_a);
        }
      return n;
    };
    std::vector<std::vector<int32_t>> groupObjs;
    std::vector<std::vector<size_t>> groupCands;
    bool single = true;
    for (size_t i = 0; i < n_cands; ++i) {
      int32_t objs[8];
      const int no = std::min(objectsOf(cands[i], objs), 8);
      size_t g = 0;
      for (; g < groupObjs.size(); ++g) {
        std::vector<int32_t> u = groupObjs[g];
        for (int k = 0; k < no; ++k) if (std::find(u.begin(), u.end(), objs[k]) == u.end()) u.push_back(objs[k]);
        if ((int)u.size() <= AFFK_MAX_OBJ || u.size() == groupObjs[g].size()) { groupObjs[g] = u; break; }
      }
      if (g == groupObjs.size()) { groupObjs.emplace_back(objs, objs + no); groupCands.emplace_back(); }
      groupCands[g].push_back(i);
      if (g > 0) single = false;
    }
    if (!single) {
      for (size_t g = 0; g < groupCands.size(); ++g) {
        std::vector<aff_candidate> local;
        for (size_t i : groupCands[g]) local.push_back(cands[i]);
        std::vector<aff_result> res(local.size());
        const int rc = aff_predict_batch(s, tasks, n_tasks, cons, n_cons, local.data(), local.size(), opts, res.data());
        if (rc != AFF_OK) return rc;
        for (size_t k = 0; k < local.size(); ++k) { out[groupCands[g][k]] = res[k]; out[groupCands[g][k]].idx = (int32_t)groupCands[g][k]; }
      }
      return AFF_OK;
    }
  }
  const auto t0 = now();
  // Large batches go through in chunks of whole waves of the persistent grid: the host compiles and
  // uploads chunk k+1 (default stream) while the kernels of chunk k run on their own stream.
  CUDA_TRY(cudaSetDevice(s->device));
  // chunk = one resident wave of the kernel the batch will run on (thread-per-candidate: 1024 rollouts per SM)
  const char* mnEnv = getenv("AFF_SWEEP_MIN");
  const char* kEnv = getenv("AFF_KERNEL");
  const bool sweepLikely = !(kEnv && !strcmp(kEnv, "warp")) && n_cands >= (mnEnv ? (size_t)atoll(mnEnv) : (size_t)AFF_SWEEP_MIN_DEFAULT);
  // (measured round 2: half-wave chunks on two alternating kernel streams, to start the first kernel earlier, lose
  // more in the kernels than they gain: e2e 99 k -> 88 k rollouts/s)
  size_t wave = sweepLikely ? (size_t)s->smCount * 1024 : (size_t)s->smCount * AFF_BLOCKS_PER_SM * AFF_WARPS_PER_BLOCK;
  // chunks in flight: at most three (running, queued, being compiled); the oldest is collected and freed
  struct Part { aff_batch* b; size_t first; cudaEvent_t done; };
  std::vector<Part> parts;
  cudaStream_t ks = nullptr, ks2 = nullptr, cs = nullptr;
  int rc = AFF_OK;
  double planMs = 0.0;
  size_t nParts = 0;
  // results of the oldest chunk -> host, then its buffers back to the pool (its kernels have finished:
  // nothing else uses them, so no further synchronisation is needed before the stream-ordered free)
  auto retire = [&](Part& p) -> int {
    int r = AFF_OK;
    const size_t cnt = (size_t)p.b->plan.nCands;
    if (p.done) {
      if (cudaEventSynchronize(p.done) != cudaSuccess) r = fail(AFF_ERR_CUDA, "aff_predict_batch: kernel failed");
      cudaEventDestroy(p.done);
      p.b->launched = false;
    }
    if (r == AFF_OK) r = aff_batch_results(p.b, cs, out + p.first, cnt);
    if (r == AFF_OK) for (size_t i = 0; i < cnt; ++i) out[p.first + i].idx += (int32_t)p.first;
    aff_batch_destroy(p.b);
    return r;
  };
  // kernels and result copies on private non-blocking streams: calls made from several host threads
  // (aff_predict_batch_multi, or one thread per action variant) overlap on the device
  if (cudaStreamCreateWithFlags(&ks, cudaStreamNonBlocking) != cudaSuccess || cudaStreamCreateWithFlags(&cs, cudaStreamNonBlocking) != cudaSuccess) {
    if (ks) cudaStreamDestroy(ks);
    ks = cs = nullptr; wave = n_cands;
  }
  if (n_cands <= wave + wave / 2) wave = n_cands;
  const size_t maxInFlight = ks2 ? 3 : 2;
  for (size_t off = 0; off < n_cands && rc == AFF_OK;) {
    size_t cnt = std::min(wave, n_cands - off);
    if (n_cands - off - cnt < wave / 2) cnt = n_cands - off;        // no short last chunk
    const auto tp = now();
    aff_batch* b = nullptr;
    if (cnt == n_cands) rc = aff_batch_create(s, tasks, n_tasks, cons, n_cons, cands, cnt, opts, &b);
    else {
      // only the slices of the task / constraint tables this chunk refers to
      int32_t t0i = INT32_MAX, t1i = 0, c0i = INT32_MAX, c1i = 0;
      for (size_t i = 0; i < cnt; ++i) {
        const aff_candidate& c = cands[off + i];
        t0i = std::min(t0i, c.first_task); t1i = std::max(t1i, c.first_task + c.n_tasks);
        c0i = std::min(c0i, c.first_constraint); c1i = std::max(c1i, c.first_constraint + c.n_constraints);
      }
      if (t0i < 0 || c0i < 0 || (size_t)t1i > n_tasks || (size_t)c1i > n_cons) { rc = fail(AFF_ERR_INVALID, "aff_predict_batch: candidate refers outside the tables"); break; }
      std::vector<aff_candidate> local(cands + off, cands + off + cnt);
      for (aff_candidate& c : local) { c.first_task -= t0i; c.first_constraint -= c0i; }
      rc = aff_batch_create(s, tasks + t0i, (size_t)(t1i - t0i), cons + c0i, (size_t)(c1i - c0i), local.data(), cnt, opts, &b);
    }
    planMs += ms(tp, now());
    if (rc != AFF_OK) break;
    Part part{b, off, nullptr};
    if (nParts == 0 && ks) {
      // the first chunk tells how many rollouts are resident at once with this batch's shared-memory footprint
      const size_t resident = b->useSweep ? (size_t)s->smCount * 1024 : (size_t)b->grid * AFF_WARPS_PER_BLOCK;
      if (resident > 0 && (resident < wave || !b->useSweep)) wave = std::min(wave, resident);
    }
    cudaStream_t kst = (ks2 && (nParts & 1)) ? ks2 : ks;
    rc = aff_batch_launch(b, kst);
    if (rc == AFF_OK && ks && (cudaEventCreateWithFlags(&part.done, cudaEventDisableTiming) != cudaSuccess || cudaEventRecord(part.done, kst) != cudaSuccess))
      rc = fail(AFF_ERR_CUDA, "aff_predict_batch: cannot record an event");
    parts.push_back(part);
    ++nParts;
    off += cnt;
    if (rc == AFF_OK && parts.size() > maxInFlight) { rc = retire(parts.front()); parts.erase(parts.begin()); }
  }
  const auto t1 = now();
  for (Part& p : parts) { const int r = retire(p); if (rc == AFF_OK) rc = r; }
  parts.clear();
  const auto t2 = now();
  if (ks) cudaStreamDestroy(ks);
  if (ks2) cudaStreamDestroy(ks2);
  if (cs) cudaStreamDestroy(cs);
  const auto t3 = now();
  if (timing) std::fprintf(stderr, "aff_predict_batch: %zu candidates in %zu chunk(s): plan + H2D %.1f ms (overlapped after the first chunk), "
                           "submit + collect %.1f ms, drain %.1f ms, streams %.1f ms\n", n_cands, nParts, planMs, ms(t0, t1), ms(t1, t2), ms(t2, t3));
  return rc;
}

int aff_predict_batch_multi(const aff_scene_desc* desc, const int* devices, int n_devices,
                            const aff_task_desc* tasks, size_t n_tasks, const aff_constraint* cons, size_t n_cons,
                            const aff_candidate* cands, size_t n_cands, const aff_opts* opts, aff_result* out)
{
  if (!desc || !devices || n_devices < 1 || !cands || !out) return fail(AFF_ERR_INVALID, "aff_predict_batch_multi: null argument");
  const int G = n_devices;
  std::vector<int> rcs((size_t)G, AFF_OK);
  std::vector<std::string> errs((size_t)G);
  std::vector<std::thread> pool;
  for (int r = 0; r < G; ++r) {
    const size_t lo = n_cands * (size_t)r / (size_t)G, hi = n_cands * (size_t)(r + 1) / (size_t)G;
    if (hi == lo) continue;
    pool.emplace_back([&, r, lo, hi]() {
      aff_scene* sc = nullptr;
      int rc = aff_scene_create(desc, devices[r], &sc);
      if (rc == AFF_OK) {
        // the slice's own window of the task / constraint tables, indices rebased
        int32_t t0 = INT32_MAX, t1 = 0, c0 = INT32_MAX, c1 = 0;
        for (size_t i = lo; i < hi; ++i) {
          t0 = std::min(t0, cands[i].first_task); t1 = std::max(t1, cands[i].first_task + cands[i].n_tasks);
          c0 = std::min(c0, cands[i].first_constraint); c1 = std::max(c1, cands[i].first_constraint + cands[i].n_constraints);
        }
        if (t0 < 0 || c0 < 0 || (size_t)t1 > n_tasks || (size_t)c1 > n_cons) { g_err = "candidate refers outside the tables"; rc = AFF_ERR_INVALID; }
        else {
          std::vector<aff_candidate> local(cands + lo, cands + hi);
          for (aff_candidate& c : local) { c.first_task -= t0; c.first_constraint -= c0; }
          rc = aff_predict_batch(sc, tasks + t0, (size_t)(t1 - t0), cons + c0, (size_t)(c1 - c0), local.data(), hi - lo, opts, out + lo);
          if (rc == AFF_OK) for (size_t i = lo; i < hi; ++i) out[i].idx += (int32_t)lo;
        }
      }
      if (rc != AFF_OK) errs[(size_t)r] = g_err;      // g_err is per thread
      aff_scene_destroy(sc);
      rcs[(size_t)r] = rc;
    });
  }
  for (std::thread& t : pool) t.join();
  for (int r = 0; r < G; ++r) if (rcs[(size_t)r] != AFF_OK) return fail(rcs[(size_t)r], "aff_predict_batch_multi: device " + std::to_string(devices[r]) + ": " + errs[(size_t)r]);
  return AFF_OK;
}

#ifdef AFFT_PROFILE
// tuning builds only: cycles thread 0 of block 0 spent in phases A..F (+ loop head) of the sweep kernel
int aff_debug_phase_clocks(unsigned long long out[16], int reset)
{
  if (cudaMemcpyFromSymbol(out, kgeneric::g_phaseClocks, sizeof(unsigned long long) * 16) != cudaSuccess) return AFF_ERR_CUDA;
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(kgeneric::g_phaseClocks, z, sizeof(z)); }
  return AFF_OK;
}
#endif

#ifdef AFFK_PROFILE
// tuning builds only: cycles lane 0 of warp 0 of block 0 spent in the parts of a step (K_MARK in rollout.cuh)
extern "C" int aff_debug_step_clocks(unsigned long long out[16], int reset)
{
  unsigned long long a[16], b[16];
  if (cudaMemcpyFromSymbol(a, kgeneric::g_kClocks, sizeof(a)) != cudaSuccess) return AFF_ERR_CUDA;
  if (cudaMemcpyFromSymbol(b, kfixed::g_kClocks, sizeof(b)) != cudaSuccess) return AFF_ERR_CUDA;
  for (int i = 0; i < 16; ++i) out[i] = a[i] + b[i];
  if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(kgeneric::g_kClocks, z, sizeof(z)); cudaMemcpyToSymbol(kfixed::g_kClocks, z, sizeof(z)); }
  return AFF_OK;
}
#endif

double aff_measure_fp64_peak(int device, double seconds)
{
  if (cudaSetDevice(device) != cudaSuccess) { g_err = "aff_measure_fp64_peak: no CUDA device"; return -1.0; }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return -1.0;
  const int threads = 256, blocks = prop.multiProcessorCount * 8, iters = 1 << 16;
  double* d = nullptr;
  if (cudaMalloc((void**)&d, sizeof(double) * threads * blocks) != cudaSuccess) return -1.0;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  aff_dfma_peak_kernel<<<blocks, threads>>>(d, 1024);   // warm-up
  cudaDeviceSynchronize();
  double best = 0.0, spent = 0.0;
  while (spent < seconds) {
    cudaEventRecord(e0);
    aff_dfma_peak_kernel<<<blocks, threads>>>(d, iters);
    cudaEventRecord(e1);
    if (cudaEventSynchronize(e1) != cudaSuccess) { best = -1.0; break; }
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double flops = 2.0 * 8.0 * (double)iters * threads * blocks;
    best = std::max(best, flops / (ms * 1e-3) * 1e-12);
    spent += ms * 1e-3 + 1e-4;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(d);
  return best;
}

}   // extern "C"
