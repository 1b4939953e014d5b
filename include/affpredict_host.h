/*
 * affpredict_host.h — C view of the host shim that sits above affpredict.h.
 *
 * The reference's front ends reach the predict step through C++ objects
 * (ActionFactory::create -> ActionBase -> predict, reference
 * src/ActionComponent.cpp:145-244) and through the pybind module
 * python/module.cpp.  These functions expose the same steps to C / ctypes:
 * load a scene, turn a text command into its Capability x Affordance
 * candidates, predict them all on the GPU and read back results + messages.
 */
#ifndef AFFPREDICT_HOST_H
#define AFFPREDICT_HOST_H

#include "affpredict.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aff_host_scene aff_host_scene;
typedef struct aff_host_batch aff_host_batch;

const char* aff_host_last_error(void);

/* Scene = kinematic graph + affordance model, from a flattened .affscene file
 * (tools/flatten_scene.py output of the reference's XML scene). NULL on error. */
aff_host_scene* aff_host_scene_load(const char* path);
void aff_host_scene_destroy(aff_host_scene* s);
/* Round trip used by the tests of the reference-tree path: rebuild the scene's graph from its own plain-C
 * description + names (Graph::fromDesc, what integration/rcs_graph_adapter.h feeds in the reference tree)
 * and report whether the rebuilt graph yields a byte-identical description.  1 = identical, 0 = not, <0 error. */
int aff_host_scene_desc_roundtrip(aff_host_scene* s);
const aff_scene_desc* aff_host_scene_desc(aff_host_scene* s);
int aff_host_num_bodies(const aff_host_scene* s);
int aff_host_num_ik_joints(const aff_host_scene* s);
int aff_host_body_id(const aff_host_scene* s, const char* name);
const char* aff_host_body_name(const aff_host_scene* s, int id);
const char* aff_host_ik_joint_name(const aff_host_scene* s, int jacobi_index);
/* Re-parent a body keeping its world pose (state left by a finished get). */
int aff_host_scene_connect(aff_host_scene* s, const char* child, const char* parent);
/* Set the unconstrained joints (jacobi order, n = aff_host_num_ik_joints) to the state a
 * finished rollout ended in: start state of the next action of a sequence. */
int aff_host_scene_set_ik_state(aff_host_scene* s, const double* q, int n);
/* Bring the scene into the state candidate `cand` of the batch leaves behind when it has run: its
 * ConnectBody / CollisionModel constraints applied at the steps they fired (with the joint state of
 * those steps), then the final joint state.  qlog = rows x nJ from aff_batch_fetch_q. */
int aff_host_scene_apply_rollout(aff_host_scene* s, const aff_host_batch* b, size_t cand, const double* qlog, int rows, double dt);
/* PredictionResult::bodyTransforms of ONE candidate (the winner, for the ghost animation of reference
 * src/ActionComponent.cpp:429-462): world transform of every body for every row of its joint log,
 * rows x aff_host_num_bodies x 12 doubles (org, then rot row-major; the layout of
 * src/TrajectoryPredictor.cpp:356-363), replayed on the host from the q(t) the GPU logged
 * (aff_batch_fetch_q; 68 MB per candidate in the pizza scene, so it is opt-in and never part of a
 * throughput run).  The scene is not modified. */
int aff_host_rollout_transforms(aff_host_scene* s, const aff_host_batch* b, size_t cand, const double* qlog, int rows,
                                double dt, double* out);
/* Device-side scene handle (created on first use). NULL on CUDA failure. */
aff_scene* aff_host_scene_device(aff_host_scene* s, int device);

aff_host_batch* aff_host_batch_create(void);
void aff_host_batch_destroy(aff_host_batch* b);
void aff_host_batch_clear(aff_host_batch* b);
/* ActionFactory::create(text) + one candidate per solution. Returns the
 * number of candidates appended, or <0 (message in aff_host_last_error —
 * the text of the reference's ActionException). */
int aff_host_batch_add_action(aff_host_batch* b, aff_host_scene* s, const char* text, double duration_scale, double dt);
/* n synthetic perturbed candidates of a get command, seeds seed0 + i. */
int aff_host_batch_add_perturbed(aff_host_batch* b, aff_host_scene* s, const char* text, size_t n, uint64_t seed0, double dt);

size_t aff_host_batch_num_tasks(const aff_host_batch* b);
size_t aff_host_batch_num_constraints(const aff_host_batch* b);
size_t aff_host_batch_num_candidates(const aff_host_batch* b);
const aff_task_desc* aff_host_batch_tasks(const aff_host_batch* b);
const aff_constraint* aff_host_batch_constraints(const aff_host_batch* b);
const aff_candidate* aff_host_batch_candidates(const aff_host_batch* b);
const char* aff_host_batch_task_name(const aff_host_batch* b, size_t cand, int task);

/* Predict every candidate of the batch on `device`; returns the count or <0. */
int aff_host_predict(aff_host_scene* s, aff_host_batch* b, double dt, int sort_results, int device,
                     aff_result* raw_out, char* messages, size_t msg_stride);
int aff_host_format_message(aff_host_scene* s, aff_host_batch* b, const aff_result* r, double dt, char* out, size_t n);
/* The same with the joint values of the failing step (row r->fail_step + 1 of the candidate's q log,
 * jacobi order): a joint-limit message then lists every violated joint with its value and range as
 * reference src/TrajectoryPredictor.cpp:810-818 does. */
int aff_host_format_message_q(aff_host_scene* s, aff_host_batch* b, const aff_result* r, double dt, const double* q_fail,
                              char* out, size_t n);

/* An action sequence "a; b; c" predicted the way the reference runs one (each action is predicted
 * from the state the previous one left, examples/ExampleActionsECS.cpp:924-979): per action all
 * candidates on the GPU, the ranked-first one is taken, its end state applied to the scene.  Stops
 * after the first action without a feasible candidate.  winners[i] / messages (msg_stride bytes each)
 * describe the ranked-first candidate of action i.  Returns the number of actions predicted, or <0.
 * The scene is left in the state after the last feasible action. */
int aff_host_predict_sequence(aff_host_scene* s, const char* text, double dt, int device,
                              aff_result* winners, int* n_candidates, char* messages, size_t msg_stride, int max_actions);

/* Second call site: TrajectoryComponent::checkerThread (reference src/TrajectoryComponent.cpp:346-399)
 * on candidate `cand` of the batch = "the runtime controller's tasks + one trajectory".
 * mode 0: onCheckAndSetTrajectory, 1: onSimulateTrajectory; estop / check_enabled are the component's
 * two switches.  Outputs: *action_result = 1 if "ActionResult" is published (then *ok, *quality and
 * message hold its arguments), *set_trajectory = 1 if "SetTrajectory" is published.  q_out (may be
 * NULL) receives up to max_rows rows of the prediction array; returns rows available or <0. */
int aff_host_check_trajectory(aff_host_scene* s, aff_host_batch* b, size_t cand, int mode, int estop, int check_enabled,
                              double dt, int device, int* action_result, int* ok, double* quality, int* set_trajectory,
                              aff_result* raw_out, char* message, size_t msg_len, double* q_out, size_t max_rows);

#ifdef __cplusplus
}
#endif
#endif
