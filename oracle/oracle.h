/*
 * oracle.h — CPU restatement of AffAction's action-feasibility predictor.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under affaction_b200/ may include, link
 * or call this; it is the checker for tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py.
 *
 * PARITY UNPINNED: the reference records no numeric output of this path and
 * its arithmetic lives in Rcs / Tropic, which are not vendored (SURVEY.md
 * 8(c)).  The loop structure, constants and checks follow the reference's own
 * src/TrajectoryPredictor.cpp line by line; the Rcs/Tropic primitives are
 * restated from their published behaviour as documented in DESIGN.md.
 *
 * The oracle consumes the same plain-C descriptions as the CUDA library
 * (include/affpredict.h) so both sides see identical inputs.
 */
#ifndef AFF_ORACLE_H
#define AFF_ORACLE_H

#include "affpredict.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Predict one candidate (ActionBase::predict + TrajectoryPredictor::predict,
 * reference src/ActionBase.cpp:116-198, src/TrajectoryPredictor.cpp:79-406).
 * `tasks` / `cons` are the batch tables the candidate slices into.
 * q_log (optional): (n_steps+1) x nJ doubles, row 0 = initial state.
 * Returns 0, or <0 on malformed input. */
int oracle_predict(const aff_scene_desc* scene,
                   const aff_task_desc* tasks, const aff_constraint* cons,
                   const aff_candidate* cand, const aff_opts* opts,
                   aff_result* out, double* q_log, size_t q_log_rows);

/* Fan a batch out over `n_threads` host threads, each candidate independent
 * (the ConcurrentExecutor path, reference src/ActionComponent.cpp:165-206). */
int oracle_predict_batch(const aff_scene_desc* scene,
                         const aff_task_desc* tasks, const aff_constraint* cons,
                         const aff_candidate* cands, size_t n_cands,
                         const aff_opts* opts, int n_threads, aff_result* out);

/* nSteps = lround(endTime/dt) + nAfterSteps (src/TrajectoryPredictor.cpp:102-103) */
int oracle_num_steps(const aff_constraint* cons, const aff_candidate* cand, const aff_opts* opts);

/* ---- building blocks exposed for property tests ---- */
/* World transforms (n_bodies*12) of the initial state (after pose override). */
int oracle_fk(const aff_scene_desc* scene, const aff_candidate* cand, double* A_BI);
/* Task value x (dim doubles) and Jacobian J (dim x nJ, row-major) at the
 * state q_ik (nJ doubles, jacobi order; NULL = initial). Returns dim. */
int oracle_task_eval(const aff_scene_desc* scene, const aff_task_desc* task,
                     const double* q_ik, double* x, double* J);
/* Task error dx for a desired value (what computeDX returns). */
int oracle_task_dx(const aff_scene_desc* scene, const aff_task_desc* task,
                   const double* q_ik, const double* x_des, double* dx);
/* Distance between two shapes given world transforms (12 doubles each). */
double oracle_shape_distance(int type1, const double* ext1, const double* A1,
                             int type2, const double* ext2, const double* A2);
/* One ViaPointTrajectory1D step: state (x,xd,xdd) advanced by dt through the
 * window of constraints (rows of t,x,xd,xdd,flag with t relative to now). */
void oracle_via_step(double state[3], const double* vias, int n_vias, double horizon, double dt);
/* Damped weighted right inverse step: dq = J#dx - (I-J#J) W^-1 dH. Returns
 * 0 if the m x m matrix is not positive definite. */
int oracle_right_inverse(int m, int n, const double* J, const double* w,
                         const double* dx, const double* dH, double lambda, double* dq);

int oracle_uses_libm(void);

#ifdef __cplusplus
}
#endif
#endif
