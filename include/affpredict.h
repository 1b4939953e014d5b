/*
 * affpredict.h — C ABI of the B200 action-feasibility predictor.
 *
 * This is the drop-in boundary for ONE path of HRI-EU/AffAction: the fan-out
 * of Capability x Affordance candidates through
 *     ActionBase::predict               (reference src/ActionBase.h:147,
 *                                        src/ActionBase.cpp:116-198)
 *     TrajectoryPredictor::predict      (src/TrajectoryPredictor.cpp:79-406)
 *     TrajectoryPredictor::computeIK    (src/TrajectoryPredictor.cpp:573-753)
 *     TrajectoryPredictor::checkState   (src/TrajectoryPredictor.cpp:756-886)
 * as driven by ActionComponent::actionThread (src/ActionComponent.cpp:157-244)
 * over a ConcurrentExecutor (src/ConcurrentExecutor.h:55-74).
 *
 * The reference has no FFI for this path (its seam is a C++ virtual), so the
 * entry points below are what a binding for it has to carry: the kinematic
 * graph the reference keeps in RcsGraph / RcsBroadPhase, the task list it
 * builds from Action*::createTasksXML, the constraint set it builds from
 * Action*::createTrajectory, and a PredictionResult per candidate.
 *
 * Plain C: pointers + sizes only, no C++ or torch types.  All floating point
 * data is IEEE double.  Transforms ("HTr", asserted to be 12 doubles at
 * src/TrajectoryPredictor.cpp:51,61) are org[3] followed by rot[9] row-major;
 * rot rows are the axes of the child frame expressed in the parent frame.
 *
 * Every function returns AFF_OK (0) or a negative aff_status; the message of
 * the last failure on the calling thread is available via aff_last_error().
 * A per-candidate infeasibility is NOT an error: it is data in aff_result,
 * exactly as TrajectoryPredictor::predict reports it (success=false+message).
 */
#ifndef AFFPREDICT_H
#define AFFPREDICT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFF_ABI_VERSION 2

/* ------------------------------------------------------------------ status */
typedef enum aff_status {
  AFF_OK = 0,
  AFF_ERR_INVALID = -1,      /* bad argument / inconsistent description      */
  AFF_ERR_CUDA = -2,         /* CUDA runtime failure (no CPU fallback exists) */
  AFF_ERR_UNSUPPORTED = -3,  /* task kind / shape pair outside this build     */
  AFF_ERR_CAPACITY = -4      /* a compile-time kernel capacity was exceeded   */
} aff_status;

/* ------------------------------------------------------------- enumerations */
/* Joint types in the order Rcs uses for rigid-body dofs. */
enum { AFF_JNT_TRANS_X = 0, AFF_JNT_TRANS_Y, AFF_JNT_TRANS_Z, AFF_JNT_ROT_X, AFF_JNT_ROT_Y, AFF_JNT_ROT_Z };

/* Distance shape primitives (Rcs SSL = sphere-swept line = capsule). */
enum { AFF_SHAPE_SSL = 0, AFF_SHAPE_SPHERE = 1, AFF_SHAPE_BOX = 2, AFF_SHAPE_CYLINDER = 3 };

/* Task kinds = Rcs controlVariable strings used by the get/put/pour actions
 * (reference src/ActionGet.cpp:848-938, src/ActionPut.cpp:379-472,
 * src/ActionPour.cpp:252-274). */
enum {
  AFF_TASK_XYZ = 0,         /* 3: position of effector w.r.t. refBdy, in refFrame */
  AFF_TASK_X = 1,           /* 1: one component of the above                      */
  AFF_TASK_Y = 2,
  AFF_TASK_Z = 3,
  AFF_TASK_POLAR = 4,       /* 2: polar angles of an effector axis in refBdy      */
  AFF_TASK_INCLINATION = 5, /* 1: angle between effector axis and refBdy z        */
  AFF_TASK_JOINTS = 6       /* n: joint angles                                    */
};

/* Constraint kinds = what a tropic::ActivationSet holds for these actions. */
enum {
  AFF_CON_VIA = 0,        /* a1->add(t, x, xd, xdd, flag, "task dim")             */
  AFF_CON_ACTIVATION = 1, /* a1->addActivation(t, on, horizon, task)               */
  AFF_CON_CONNECT = 2,    /* tropic::ConnectBodyConstraint(t, child, parent); dim = 1:
                           * setConnectTransform(HTr_identity()), the child jumps onto the parent frame */
  AFF_CON_COLLISION = 3   /* tropic::CollisionModelConstraint(t, body, on)         */
};

/* computeIK return codes (reference src/TrajectoryPredictor.h:150-151) plus
 * the two failure kinds TrajectoryPredictor::predict raises itself. */
enum {
  AFF_CODE_OK = 0,
  AFF_CODE_SINGULAR = -1,
  AFF_CODE_SPEED = -2,
  AFF_CODE_JOINT_LIMIT = -3,
  AFF_CODE_COLLISION = -4,
  AFF_CODE_TRACKING = -5,      /* tracking error > 0.05 (src/TrajectoryPredictor.cpp:304) */
  AFF_CODE_INITIAL_STATE = -6  /* check() failed before the loop (:135-140)               */
};

#define AFF_MAX_TASK_JOINTS 8

/* ------------------------------------------------------------------- scene */
/* Flattened RcsGraph + RcsBroadPhase.  Joints of one body are contiguous and
 * applied in order, then the body's own A_BP (Rcs convention).  Parents always
 * have a smaller index than their children at creation time. */
typedef struct aff_scene_desc {
  int32_t n_bodies, n_joints, n_shapes;

  const int32_t* body_parent;      /* [n_bodies]  -1 = world                      */
  const int32_t* body_first_joint; /* [n_bodies]                                  */
  const int32_t* body_n_joints;    /* [n_bodies]                                  */
  const int32_t* body_first_shape; /* [n_bodies]                                  */
  const int32_t* body_n_shapes;    /* [n_bodies]                                  */
  const uint8_t* body_rbj;         /* [n_bodies]  1 = has six rigid-body joints   */
  const double* body_A_BP;         /* [n_bodies*12]                               */

  const int32_t* joint_type;       /* [n_joints]  AFF_JNT_*                       */
  const uint8_t* joint_constrained;/* [n_joints]  1 = not an IK dof               */
  const double* joint_A_JP;        /* [n_joints*12]                               */
  const double* joint_q_init;      /* [n_joints]  state the prediction starts at  */
  const double* joint_q0;          /* [n_joints]  centre of range (JL cost)       */
  const double* joint_q_min;       /* [n_joints]                                  */
  const double* joint_q_max;       /* [n_joints]                                  */
  const double* joint_speed_limit; /* [n_joints]  rad/s or m/s; >=1e300 = none    */
  const double* joint_acc_limit;   /* [n_joints]                                  */
  const double* joint_weight_jl;   /* [n_joints]                                  */
  const double* joint_weight_metric;/*[n_joints]  Rcs invWq entry                 */

  const int32_t* shape_type;       /* [n_shapes]  AFF_SHAPE_*                     */
  const uint8_t* shape_distance;   /* [n_shapes]  RCSSHAPE_COMPUTE_DISTANCE bit   */
  const double* shape_A_CB;        /* [n_shapes*12]                               */
  const double* shape_extents;     /* [n_shapes*3] BOX: xyz; others: r, r, length */

  double bp_threshold;             /* <BroadPhase DistanceThreshold>              */
  int32_t n_bp_trees;  const int32_t* bp_tree_root; /* <Tree root>               */
  int32_t n_bp_bodies; const int32_t* bp_body;      /* <Body name>               */

  /* Bodies the elbow / wrist null-space terms look up by name
   * (src/TrajectoryPredictor.cpp:459-463,492-493,540-542); -1 = absent.
   * Index 0 = left, 1 = right. */
  int32_t ns_base, ns_upperarm[2], ns_forearm[2], ns_hand[2];
} aff_scene_desc;

/* -------------------------------------------------------------------- task */
typedef struct aff_task_desc {
  int32_t kind;       /* AFF_TASK_*                                              */
  int32_t effector;   /* body id                                                 */
  int32_t ref_body;   /* body id or -1 (world)                                   */
  int32_t ref_frame;  /* body id, -1 (= ref_body), or -2 (= world frame whatever ref_body is) */
  int32_t axis;       /* POLAR / INCLINATION axisDirection: 0,1,2 (default 2)    */
  int32_t n_joints;   /* JOINTS only                                             */
  int32_t joints[AFF_MAX_TASK_JOINTS]; /* joint ids                              */
  int32_t has_region; /* <TaskRegion type="BoxInterval">                         */
  double region_min, region_max, region_dx_scaling;
} aff_task_desc;

/* -------------------------------------------------------------- constraint */
typedef struct aff_constraint {
  int32_t kind;       /* AFF_CON_*                                               */
  int32_t task;       /* VIA / ACTIVATION: index into the candidate's task list  */
  int32_t dim;        /* VIA: task dimension                                     */
  int32_t flag;       /* VIA: bit0 pos, bit1 vel, bit2 acc; others: 1 = on       */
  int32_t body_a;     /* CONNECT: child;  COLLISION: body                        */
  int32_t body_b;     /* CONNECT: new parent                                     */
  double t;           /* seconds from prediction start                           */
  double x, xd, xdd;  /* VIA                                                     */
  double horizon;     /* ACTIVATION: blending horizon (0.5 in all shipped uses)  */
} aff_constraint;

/* --------------------------------------------------------------- candidate */
/* One Capability x Affordance solution = one rollout: its own task list and
 * constraint set (ActionBase::predict builds both per candidate), plus an
 * optional pose override of one rigid-body-joint body (synthetic sweeps
 * perturb the object pose per rollout). */
typedef struct aff_candidate {
  int32_t first_task, n_tasks;             /* slice of the batch task table     */
  int32_t first_constraint, n_constraints; /* slice of the batch constraint tab */
  int32_t pose_body;                       /* -1 or a body with rbj             */
  int32_t reserved;
  double pose_q6[6];                       /* its six rigid-body dof values     */
} aff_candidate;

/* ------------------------------------------------------------------ result */
/* PredictionResult without its strings and its 68 MB transform log
 * (reference src/TrajectoryPredictor.h:48-124).  The host shim rebuilds the
 * message strings from (code, fail_step, fail_id0, fail_id1). */
typedef struct aff_result {
  int32_t idx;         /* candidate index within the batch                       */
  int32_t success;     /* 1 / 0                                                  */
  int32_t code;        /* AFF_CODE_* of the LAST failing step (message is        */
                       /* overwritten each failing step, :220-224)               */
  int32_t first_code;  /* AFF_CODE_* of the first failing step                   */
  int32_t first_fail_step; /* step index of the first failure, -1 if none        */
  int32_t fail_step;   /* step index of the last failure                         */
  int32_t fail_id0;    /* last failure: pair body 1 / #violated joints / task    */
  int32_t fail_id1;    /* last failure: pair body 2 / task dimension             */
  int32_t min_dist_b1; /* bodies of the closest pair over the rollout, -1 = NULL */
  int32_t min_dist_b2;
  int32_t n_steps;     /* rows accumulated = nSteps+1 (0 if initial check failed)*/
  uint32_t reserved;        /* flags: AFF_FLAG_LIST_OVERFLOW */
  uint64_t jmask_bits; /* binarised jMask, bit = jacobi index                    */
  double min_dist;     /* DBL_MAX if no pair was ever tested                     */
  double jl_cost;      /* mean over rows                                         */
  double coll_cost;    /* mean over rows                                         */
  double track_err;    /* running max tracking error while successful            */
} aff_result;

/* aff_result.reserved, bit 0: in some step a scratch list of the collision model overflowed (more than 32
 * body pairs closer than DistanceThreshold, or more than 64 body pairs found by the AABB test): the entries
 * beyond the capacity were dropped, so coll_cost / min_dist of this candidate may be incomplete.  Never seen
 * in the shipped scenes; reported instead of truncating silently. */
#define AFF_FLAG_LIST_OVERFLOW 1u

typedef struct aff_opts {
  double dt;            /* 0.01 in every reference call site                     */
  double alpha;         /* 0.05  (src/TrajectoryPredictor.cpp:149)               */
  double lambda;        /* 1e-6  (:150)                                          */
  double q_filt_decay;  /* 0.1   (:151)                                          */
  int32_t n_after_steps;/* 500   (:102)                                          */
  int32_t log_q;        /* 1: keep q(t) for aff_batch_fetch_q (parity samples)   */
  int32_t early_exit;   /* 0 (default, the reference's behaviour: every rollout runs all its steps).
                           1: a rollout stops at the end of the step in which it first failed - the
                           `break` the reference keeps commented out (:368-373).  Verdict, failure
                           code / step / ids are unchanged; the costs of FAILED candidates are then
                           averaged over fewer steps, so their order among themselves may differ. */
  int32_t reserved;
} aff_opts;

void aff_default_opts(aff_opts* o);

/* ------------------------------------------------------------------ handles */
typedef struct aff_scene aff_scene;
typedef struct aff_batch aff_batch;

int aff_abi_version(void);
const char* aff_last_error(void);

/* Number of CUDA devices visible; <0 on CUDA failure. */
int aff_device_count(void);

/* Copies the description; the caller's arrays may be freed after return.
 * Replaces: RcsGraph_clone + RcsBroadPhase_clone + RcsCollisionModel_create
 * per candidate (src/ActionBase.cpp:123-131) by one upload per scene. */
int aff_scene_create(const aff_scene_desc* desc, int device, aff_scene** out);
void aff_scene_destroy(aff_scene* s);
int aff_scene_num_ik_joints(const aff_scene* s);   /* RcsGraph::nJ             */

/* Compiles and uploads a batch (tasks + constraints + candidates): inputs
 * become resident in HBM.  Replaces addTasks / createTrajectory /
 * TrajectoryPredictor ctor + setTrajectory (src/ActionBase.cpp:136-141). */
int aff_batch_create(aff_scene* s,
                     const aff_task_desc* tasks, size_t n_tasks,
                     const aff_constraint* cons, size_t n_cons,
                     const aff_candidate* cands, size_t n_cands,
                     const aff_opts* opts, aff_batch** out);
void aff_batch_destroy(aff_batch* b);

/* Launches all rollouts of the batch on `stream` (a cudaStream_t, or NULL for
 * the default stream).  Asynchronous.  Results stay on the device.
 * Replaces the ConcurrentExecutor fan-out (src/ActionComponent.cpp:172-199). */
int aff_batch_launch(aff_batch* b, void* stream);

/* Device pointer of the n_cands aff_result records (valid until destroy). */
void* aff_batch_results_device(aff_batch* b);
/* Make the kernel write its n_cands records into caller-owned device memory
 * (e.g. the send buffer of an NCCL gather); NULL restores the internal buffer. */
int aff_batch_set_results_buffer(aff_batch* b, void* device_ptr);
/* Blocking copy of the results to host memory (the future.wait() join,
 * src/ActionComponent.cpp:202-205). */
int aff_batch_results(aff_batch* b, void* stream, aff_result* out, size_t n);

/* Kernels launched by the last aff_batch_launch (bench gpu_launches claim). */
int aff_batch_last_launch_count(const aff_batch* b);
/* Steps a full-length rollout of candidate i runs (nSteps). */
int aff_batch_num_steps(const aff_batch* b, size_t cand);

/* Bytes copied host->device when the batch was created, shared-memory bytes one
 * rollout (warp) uses, and the persistent grid size. */
size_t aff_batch_h2d_bytes(const aff_batch* b);
size_t aff_batch_smem_bytes_per_warp(const aff_batch* b);
int aff_batch_grid_blocks(const aff_batch* b);

/* q(t) log of one candidate: (n_steps+1) rows of nJ doubles, jacobi-index
 * order; requires opts.log_q.  Returns rows written or <0. */
int aff_batch_fetch_q(aff_batch* b, size_t cand, double* out, size_t max_rows);

/* One-call form with HOST buffers in and out: upload, launch, download.
 * This is the call the reference-facing shim makes per action. */
int aff_predict_batch(aff_scene* s,
                      const aff_task_desc* tasks, size_t n_tasks,
                      const aff_constraint* cons, size_t n_cons,
                      const aff_candidate* cands, size_t n_cands,
                      const aff_opts* opts, aff_result* out);

/* The same over several GPUs of one box from ONE host process (the reference's
 * ActionComponent::actionThread is one process, src/ActionComponent.cpp:163-244):
 * candidates are split statically, device r of G takes [r n/G, (r+1) n/G)
 * (independent units: no data-path collective), each slice runs through
 * aff_predict_batch on its own host thread and device, and the records land in
 * the caller's array in candidate order.  `devices` may name a device twice.
 * Returns AFF_OK or the first failing slice's status. */
int aff_predict_batch_multi(const aff_scene_desc* scene, const int* devices, int n_devices,
                            const aff_task_desc* tasks, size_t n_tasks,
                            const aff_constraint* cons, size_t n_cons,
                            const aff_candidate* cands, size_t n_cands,
                            const aff_opts* opts, aff_result* out);

/* Which kernel a batch runs on: 0 = one rollout per warp (mid-size batches),
 * 1 = one rollout per thread (sweeps), 2 = one rollout per block (a handful of
 * candidates, one trajectory re-check); see DESIGN.md section 4.  The records
 * are the same bit for bit whichever runs. */
int aff_batch_kernel_kind(const aff_batch* b);

/* Argsort of results with PredictionResult::lesser (success first, then
 * jlCost+collCost ascending; src/TrajectoryPredictor.h:61-74), ties broken
 * by idx so the order is deterministic (std::sort in the reference is not). */
void aff_rank_results(const aff_result* res, size_t n, int32_t* order);

/* FP64 FMA throughput microbenchmark (register-resident DFMA chains) used as
 * the roofline denominator; returns TFLOP/s or <0. */
double aff_measure_fp64_peak(int device, double seconds);

#ifdef __cplusplus
}
#endif
#endif /* AFFPREDICT_H */
