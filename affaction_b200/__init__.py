"""affaction_b200 — ctypes binding of the B200 action-feasibility predictor.

The product is two shared libraries: ``libaffpredict.so`` (hand-written sm_100a
CUDA kernels behind the C ABI of ``include/affpredict.h``) and ``libaffhost.so``
(the C++ host shim of ``include/affpredict_host.h``: pure C++, it reaches the CUDA
library through dlopen when something is predicted).  This module only binds it for the
tests and for ``bench.py``; there is no Python or CPU implementation of the
predict path here, and importing the library fails loudly if it has not been
built (``make`` / ``__graft_entry__.build()``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
REPO_ROOT = os.path.dirname(_HERE)
# AFF_LIB_PATH: tuning experiments only (tools/try_variants.sh builds alternative libraries)
LIB_PATH = os.environ.get("AFF_LIB_PATH") or os.path.join(_HERE, "libaffpredict.so")
HOST_LIB_PATH = os.path.join(_HERE, "libaffhost.so")
DATA_DIR = os.path.join(_HERE, "data")
PIZZA_SCENE = os.path.join(DATA_DIR, "pizza.affscene")

AFF_OK = 0
AFF_MAX_TASK_JOINTS = 8

# enums of include/affpredict.h
TASK_XYZ, TASK_X, TASK_Y, TASK_Z, TASK_POLAR, TASK_INCLINATION, TASK_JOINTS = range(7)
CON_VIA, CON_ACTIVATION, CON_CONNECT, CON_COLLISION = range(4)
CODE_NAMES = {0: "ok", -1: "singular", -2: "speed", -3: "joint_limit", -4: "collision", -5: "tracking", -6: "initial_state"}


class Result(C.Structure):
    _fields_ = [("idx", C.c_int32), ("success", C.c_int32), ("code", C.c_int32), ("first_code", C.c_int32),
                ("first_fail_step", C.c_int32), ("fail_step", C.c_int32), ("fail_id0", C.c_int32), ("fail_id1", C.c_int32),
                ("min_dist_b1", C.c_int32), ("min_dist_b2", C.c_int32), ("n_steps", C.c_int32), ("reserved", C.c_uint32),
                ("jmask_bits", C.c_uint64), ("min_dist", C.c_double), ("jl_cost", C.c_double), ("coll_cost", C.c_double),
                ("track_err", C.c_double)]

    COMPARE_FIELDS = ("success", "code", "first_code", "first_fail_step", "fail_step", "fail_id0", "fail_id1", "min_dist_b1",
                      "min_dist_b2", "n_steps", "jmask_bits", "min_dist", "jl_cost", "coll_cost", "track_err")

    def as_tuple(self):
        return tuple(getattr(self, f) for f in self.COMPARE_FIELDS)


class Opts(C.Structure):
    _fields_ = [("dt", C.c_double), ("alpha", C.c_double), ("lam", C.c_double), ("q_filt_decay", C.c_double),
                ("n_after_steps", C.c_int32), ("log_q", C.c_int32), ("early_exit", C.c_int32), ("reserved", C.c_int32)]


class TaskDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("effector", C.c_int32), ("ref_body", C.c_int32), ("ref_frame", C.c_int32),
                ("axis", C.c_int32), ("n_joints", C.c_int32), ("joints", C.c_int32 * AFF_MAX_TASK_JOINTS),
                ("has_region", C.c_int32), ("region_min", C.c_double), ("region_max", C.c_double),
                ("region_dx_scaling", C.c_double)]


class Constraint(C.Structure):
    _fields_ = [("kind", C.c_int32), ("task", C.c_int32), ("dim", C.c_int32), ("flag", C.c_int32), ("body_a", C.c_int32),
                ("body_b", C.c_int32), ("t", C.c_double), ("x", C.c_double), ("xd", C.c_double), ("xdd", C.c_double),
                ("horizon", C.c_double)]


class Candidate(C.Structure):
    _fields_ = [("first_task", C.c_int32), ("n_tasks", C.c_int32), ("first_constraint", C.c_int32),
                ("n_constraints", C.c_int32), ("pose_body", C.c_int32), ("reserved", C.c_int32), ("pose_q6", C.c_double * 6)]


def default_opts(log_q=False, early_exit=False):
    o = Opts()
    lib().aff_default_opts(C.byref(o))
    o.log_q = 1 if log_q else 0
    o.early_exit = 1 if early_exit else 0
    return o


# every symbol include/affpredict.h and include/affpredict_host.h declare
ABI_SYMBOLS = [
    "aff_default_opts", "aff_abi_version", "aff_last_error", "aff_device_count", "aff_scene_create", "aff_scene_destroy",
    "aff_scene_num_ik_joints", "aff_batch_create", "aff_batch_destroy", "aff_batch_launch", "aff_batch_results_device",
    "aff_batch_results", "aff_batch_last_launch_count", "aff_batch_num_steps", "aff_batch_fetch_q", "aff_predict_batch",
    "aff_predict_batch_multi", "aff_batch_kernel_kind", "aff_rank_results", "aff_measure_fp64_peak", "aff_batch_set_results_buffer", "aff_batch_h2d_bytes", "aff_batch_smem_bytes_per_warp", "aff_batch_grid_blocks",
]
HOST_SYMBOLS = [
    "aff_host_last_error", "aff_host_scene_load", "aff_host_scene_destroy", "aff_host_scene_desc", "aff_host_num_bodies",
    "aff_host_num_ik_joints", "aff_host_body_id", "aff_host_body_name", "aff_host_ik_joint_name", "aff_host_scene_connect", "aff_host_scene_set_ik_state",
    "aff_host_scene_device", "aff_host_batch_create", "aff_host_batch_destroy", "aff_host_batch_clear",
    "aff_host_batch_add_action", "aff_host_batch_add_perturbed", "aff_host_batch_num_tasks", "aff_host_batch_num_constraints",
    "aff_host_batch_num_candidates", "aff_host_batch_tasks", "aff_host_batch_constraints", "aff_host_batch_candidates",
    "aff_host_batch_task_name", "aff_host_predict", "aff_host_format_message", "aff_host_check_trajectory", "aff_host_scene_apply_rollout", "aff_host_predict_sequence", "aff_host_rollout_transforms", "aff_host_format_message_q", "aff_host_scene_desc_roundtrip",
]


class _Libs:
    """Attribute lookup over the two libraries: host-shim symbols come from libaffhost.so, the rest from
    libaffpredict.so (loaded on first use of one of ITS symbols, so that code which only builds candidates -
    the CPU reference arm of bench.py - never maps the CUDA library)."""

    def __init__(self):
        self._host = None
        self._gpu = None

    def host(self):
        if self._host is None:
            if not os.path.isfile(HOST_LIB_PATH):
                raise RuntimeError("%s is missing: build it with `make`" % HOST_LIB_PATH)
            self._host = C.CDLL(HOST_LIB_PATH)
            _declare(self._host, HOST_SYMBOLS + ["aff_default_opts", "aff_abi_version", "aff_rank_results"])
        return self._host

    def gpu(self):
        if self._gpu is None:
            if not os.path.isfile(LIB_PATH):
                raise RuntimeError("%s is missing: build it with `make` (there is no CPU fallback for the predict path)" % LIB_PATH)
            self._gpu = C.CDLL(LIB_PATH)
            _declare(self._gpu, ABI_SYMBOLS)
        return self._gpu

    def __getattr__(self, name):
        if name.startswith("aff_host_") or name in ("aff_default_opts", "aff_abi_version", "aff_rank_results"):
            return getattr(self.host(), name)
        return getattr(self.gpu(), name)


def _signatures():
    vp, sz, dbl, i32, cstr = C.c_void_p, C.c_size_t, C.c_double, C.c_int, C.c_char_p
    return {
        "aff_last_error": (cstr, []), "aff_abi_version": (i32, []), "aff_device_count": (i32, []),
        "aff_default_opts": (None, [vp]),
        "aff_scene_create": (i32, [vp, i32, C.POINTER(vp)]), "aff_scene_destroy": (None, [vp]),
        "aff_scene_num_ik_joints": (i32, [vp]),
        "aff_batch_create": (i32, [vp, vp, sz, vp, sz, vp, sz, vp, C.POINTER(vp)]), "aff_batch_destroy": (None, [vp]),
        "aff_batch_launch": (i32, [vp, vp]), "aff_batch_results_device": (vp, [vp]),
        "aff_batch_results": (i32, [vp, vp, vp, sz]), "aff_batch_set_results_buffer": (i32, [vp, vp]), "aff_batch_last_launch_count": (i32, [vp]),
        "aff_batch_num_steps": (i32, [vp, sz]), "aff_batch_fetch_q": (i32, [vp, sz, vp, sz]),
        "aff_predict_batch": (i32, [vp, vp, sz, vp, sz, vp, sz, vp, vp]),
        "aff_predict_batch_multi": (i32, [vp, vp, i32, vp, sz, vp, sz, vp, sz, vp, vp]), "aff_batch_kernel_kind": (i32, [vp]), "aff_rank_results": (None, [vp, sz, vp]),
        "aff_measure_fp64_peak": (dbl, [i32, dbl]), "aff_batch_h2d_bytes": (sz, [vp]),
        "aff_batch_smem_bytes_per_warp": (sz, [vp]), "aff_batch_grid_blocks": (i32, [vp]),
        "aff_host_last_error": (cstr, []), "aff_host_scene_load": (vp, [cstr]), "aff_host_scene_destroy": (None, [vp]),
        "aff_host_scene_desc": (vp, [vp]), "aff_host_num_bodies": (i32, [vp]), "aff_host_num_ik_joints": (i32, [vp]),
        "aff_host_body_id": (i32, [vp, cstr]), "aff_host_body_name": (cstr, [vp, i32]),
        "aff_host_ik_joint_name": (cstr, [vp, i32]), "aff_host_scene_connect": (i32, [vp, cstr, cstr]), "aff_host_scene_set_ik_state": (i32, [vp, vp, i32]),
        "aff_host_scene_device": (vp, [vp, i32]), "aff_host_batch_create": (vp, []), "aff_host_batch_destroy": (None, [vp]),
        "aff_host_batch_clear": (None, [vp]), "aff_host_batch_add_action": (i32, [vp, vp, cstr, dbl, dbl]),
        "aff_host_batch_add_perturbed": (i32, [vp, vp, cstr, sz, C.c_uint64, dbl]),
        "aff_host_batch_num_tasks": (sz, [vp]), "aff_host_batch_num_constraints": (sz, [vp]),
        "aff_host_batch_num_candidates": (sz, [vp]), "aff_host_batch_tasks": (vp, [vp]),
        "aff_host_batch_constraints": (vp, [vp]), "aff_host_batch_candidates": (vp, [vp]),
        "aff_host_batch_task_name": (cstr, [vp, sz, i32]),
        "aff_host_predict": (i32, [vp, vp, dbl, i32, i32, vp, vp, sz]),
        "aff_host_format_message": (i32, [vp, vp, vp, dbl, vp, sz]),
        "aff_host_scene_apply_rollout": (i32, [vp, vp, sz, vp, i32, dbl]),
        "aff_host_rollout_transforms": (i32, [vp, vp, sz, vp, i32, dbl, vp]),
        "aff_host_scene_desc_roundtrip": (i32, [vp]),
        "aff_host_format_message_q": (i32, [vp, vp, vp, dbl, vp, vp, sz]),
        "aff_host_predict_sequence": (i32, [vp, cstr, dbl, i32, vp, vp, vp, sz, i32]),
        "aff_host_check_trajectory": (i32, [vp, vp, sz, i32, i32, i32, dbl, i32, vp, vp, vp, vp, vp, vp, sz, vp, sz]),
    }


def _declare(L, names):
    sig = _signatures()
    for name in names:
        if name not in sig:
            continue
        fn = getattr(L, name)
        fn.restype, fn.argtypes = sig[name]


_libs = _Libs()


def lib():
    """Both libraries behind one attribute lookup (see _Libs).  Raises if a library is missing."""
    return _libs


def gpu_library_loaded():
    """True once libaffpredict.so has been mapped by this module (the CPU reference arm checks that it never is)."""
    return _libs._gpu is not None


class AffError(RuntimeError):
    pass


def _check(rc, what):
    if rc != AFF_OK:
        raise AffError("%s failed (%d): %s" % (what, rc, lib().aff_last_error().decode()))


class HostScene:
    """Kinematic graph + affordance model loaded from a flattened .affscene file."""

    def __init__(self, path=PIZZA_SCENE):
        self.h = lib().aff_host_scene_load(path.encode())
        if not self.h:
            raise AffError("cannot load scene %s: %s" % (path, lib().aff_host_last_error().decode()))
        self.path = path

    def close(self):
        if self.h:
            lib().aff_host_scene_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def desc(self):
        return lib().aff_host_scene_desc(self.h)

    @property
    def n_bodies(self):
        return lib().aff_host_num_bodies(self.h)

    @property
    def nJ(self):
        return lib().aff_host_num_ik_joints(self.h)

    def body_id(self, name):
        return lib().aff_host_body_id(self.h, name.encode())

    def body_name(self, i):
        return lib().aff_host_body_name(self.h, i).decode()

    def joint_name(self, c):
        return lib().aff_host_ik_joint_name(self.h, c).decode()

    def connect(self, child, parent):
        rc = lib().aff_host_scene_connect(self.h, child.encode(), parent.encode())
        if rc != AFF_OK:
            raise AffError("connect failed: %s" % lib().aff_host_last_error().decode())

    def set_ik_state(self, q):
        import numpy as np
        q = np.ascontiguousarray(q, dtype=np.float64)
        rc = lib().aff_host_scene_set_ik_state(self.h, q.ctypes.data_as(C.c_void_p), len(q))
        if rc != AFF_OK:
            raise AffError("set_ik_state failed: %s" % lib().aff_host_last_error().decode())

    def apply_rollout(self, batch, cand, q, dt=0.01):
        """Scene state after candidate `cand` of `batch` has run (q = its joint log, rows x nJ)."""
        import numpy as np
        q = np.ascontiguousarray(q, dtype=np.float64)
        rc = lib().aff_host_scene_apply_rollout(self.h, batch.h, cand, q.ctypes.data_as(C.c_void_p), q.shape[0], dt)
        if rc != AFF_OK:
            raise AffError("apply_rollout failed: %s" % lib().aff_host_last_error().decode())

    def rollout_transforms(self, batch, cand, q, dt=0.01):
        """PredictionResult::bodyTransforms of one candidate: rows x n_bodies x 12 from its joint log q."""
        import numpy as np
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.zeros((q.shape[0], self.n_bodies, 12))
        rc = lib().aff_host_rollout_transforms(self.h, batch.h, cand, q.ctypes.data_as(C.c_void_p), q.shape[0], dt,
                                               out.ctypes.data_as(C.c_void_p))
        if rc != AFF_OK:
            raise AffError("rollout_transforms failed: %s" % lib().aff_host_last_error().decode())
        return out

    def predict_sequence(self, text, dt=0.01, device=0, max_actions=16, msg_len=768):
        """aff_host_predict_sequence: [(winner record, number of candidates, message)] per action."""
        res = (Result * max_actions)()
        ncand = (C.c_int * max_actions)()
        msgs = C.create_string_buffer(max_actions * msg_len)
        n = lib().aff_host_predict_sequence(self.h, text.encode(), dt, device, res, ncand, msgs, msg_len, max_actions)
        if n < 0:
            raise AffError("predict_sequence: %s / %s" % (lib().aff_host_last_error().decode(), lib().aff_last_error().decode()))
        return [(res[i], ncand[i], msgs.raw[i * msg_len:(i + 1) * msg_len].split(b"\0", 1)[0].decode()) for i in range(n)]

    def device_scene(self, device=0):
        p = lib().aff_host_scene_device(self.h, device)
        if not p:
            raise AffError("no device scene: %s" % lib().aff_host_last_error().decode())
        return p


class HostBatch:
    """Candidates (tasks + constraints) generated from text commands."""

    def __init__(self, scene):
        self.scene = scene
        self.h = lib().aff_host_batch_create()

    def close(self):
        if self.h:
            lib().aff_host_batch_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def clear(self):
        lib().aff_host_batch_clear(self.h)

    def add_action(self, text, duration_scale=1.0, dt=0.01):
        n = lib().aff_host_batch_add_action(self.h, self.scene.h, text.encode(), duration_scale, dt)
        if n < 0:
            raise AffError(lib().aff_host_last_error().decode())
        return n

    def add_perturbed(self, text, n, seed0=0xAFFAC7, dt=0.01):
        r = lib().aff_host_batch_add_perturbed(self.h, self.scene.h, text.encode(), n, seed0, dt)
        if r < 0:
            raise AffError(lib().aff_host_last_error().decode())
        return r

    @property
    def n_tasks(self):
        return lib().aff_host_batch_num_tasks(self.h)

    @property
    def n_constraints(self):
        return lib().aff_host_batch_num_constraints(self.h)

    @property
    def n_candidates(self):
        return lib().aff_host_batch_num_candidates(self.h)

    @property
    def tasks_ptr(self):
        return lib().aff_host_batch_tasks(self.h)

    @property
    def constraints_ptr(self):
        return lib().aff_host_batch_constraints(self.h)

    @property
    def candidates_ptr(self):
        return lib().aff_host_batch_candidates(self.h)

    def candidate_ptr(self, i):
        return self.candidates_ptr + i * C.sizeof(Candidate)

    def task_name(self, cand, task):
        return lib().aff_host_batch_task_name(self.h, cand, task).decode()

    def host_bytes(self):
        return (self.n_tasks * C.sizeof(TaskDesc) + self.n_constraints * C.sizeof(Constraint) +
                self.n_candidates * C.sizeof(Candidate))

    def predict(self, dt=0.01, sort=False, device=0, msg_len=768):
        """aff_host_predict: all candidates on the GPU -> (results, messages)."""
        n = self.n_candidates
        res = (Result * n)()
        msgs = C.create_string_buffer(n * msg_len)
        rc = lib().aff_host_predict(self.scene.h, self.h, dt, 1 if sort else 0, device, res, msgs, msg_len)
        if rc < 0:
            raise AffError("aff_host_predict: %s / %s" % (lib().aff_host_last_error().decode(), lib().aff_last_error().decode()))
        out = [msgs.raw[i * msg_len:(i + 1) * msg_len].split(b"\0", 1)[0].decode() for i in range(n)]
        return list(res), out

    def check_trajectory(self, cand=0, simulate_only=False, estop=False, check_enabled=True, dt=0.01, device=0, max_rows=4096):
        """TrajectoryComponent::onCheckAndSetTrajectory / onSimulateTrajectory on one candidate.
        Returns a dict with the events the component publishes and the prediction array."""
        import numpy as np
        nJ = self.scene.nJ
        q = np.zeros((max_rows, nJ))
        ar, ok, st = C.c_int(0), C.c_int(0), C.c_int(0)
        quality = C.c_double(0.0)
        raw = Result()
        msg = C.create_string_buffer(1024)
        rows = lib().aff_host_check_trajectory(self.scene.h, self.h, cand, 1 if simulate_only else 0, 1 if estop else 0,
                                               1 if check_enabled else 0, dt, device, C.byref(ar), C.byref(ok), C.byref(quality),
                                               C.byref(st), C.byref(raw), msg, 1024, q.ctypes.data, max_rows)
        if rows < 0:
            raise AffError("aff_host_check_trajectory: %s / %s" % (lib().aff_host_last_error().decode(), lib().aff_last_error().decode()))
        return {"action_result": bool(ar.value), "ok": bool(ok.value), "quality": quality.value, "set_trajectory": bool(st.value),
                "message": msg.value.decode(), "raw": raw, "q": q[:min(rows, max_rows)].copy()}

    def format_message(self, result, dt=0.01, q_fail=None):
        """q_fail: joint values of the failing step (row fail_step + 1 of the q log) for the per-joint detail."""
        buf = C.create_string_buffer(4096)
        if q_fail is None:
            lib().aff_host_format_message(self.scene.h, self.h, C.byref(result), dt, buf, 4096)
        else:
            import numpy as np
            qf = np.ascontiguousarray(q_fail, dtype=np.float64)
            lib().aff_host_format_message_q(self.scene.h, self.h, C.byref(result), dt, qf.ctypes.data_as(C.c_void_p), buf, 4096)
        return buf.value.decode()


class DeviceBatch:
    """aff_batch: a candidate batch resident on one GPU."""

    def __init__(self, scene, batch, opts=None, device=0):
        self.scene, self.batch = scene, batch
        self.opts = opts if opts is not None else default_opts()
        self.h = C.c_void_p()
        _check(lib().aff_batch_create(scene.device_scene(device), batch.tasks_ptr, batch.n_tasks, batch.constraints_ptr,
                                      batch.n_constraints, batch.candidates_ptr, batch.n_candidates, C.byref(self.opts),
                                      C.byref(self.h)), "aff_batch_create")
        self.n = batch.n_candidates

    def close(self):
        if self.h:
            lib().aff_batch_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launch(self, stream=None):
        _check(lib().aff_batch_launch(self.h, stream), "aff_batch_launch")
        return lib().aff_batch_last_launch_count(self.h)

    def results(self, stream=None):
        res = (Result * self.n)()
        _check(lib().aff_batch_results(self.h, stream, res, self.n), "aff_batch_results")
        return list(res)

    def set_results_buffer(self, device_ptr):
        _check(lib().aff_batch_set_results_buffer(self.h, device_ptr), "aff_batch_set_results_buffer")

    @property
    def results_device_ptr(self):
        return lib().aff_batch_results_device(self.h)

    def num_steps(self, cand=0):
        return lib().aff_batch_num_steps(self.h, cand)

    def fetch_q(self, cand):
        import numpy as np
        rows = self.num_steps(cand) + 1
        out = np.zeros((rows, self.scene.nJ))
        r = lib().aff_batch_fetch_q(self.h, cand, out.ctypes.data_as(C.c_void_p), rows)
        if r < 0:
            raise AffError(lib().aff_last_error().decode())
        return out[:r]

    @property
    def h2d_bytes(self):
        return lib().aff_batch_h2d_bytes(self.h)

    @property
    def smem_bytes_per_warp(self):
        return lib().aff_batch_smem_bytes_per_warp(self.h)

    @property
    def grid_blocks(self):
        return lib().aff_batch_grid_blocks(self.h)

    @property
    def kernel(self):
        """'cta' (one rollout per block: a handful of candidates), 'warp' (one rollout per warp) or
        'sweep' (one rollout per thread)."""
        return ("warp", "sweep", "cta")[lib().aff_batch_kernel_kind(self.h)]


def predict_batch(scene, batch, opts=None, device=0):
    """aff_predict_batch: host buffers in, host results out (the e2e call)."""
    opts = opts if opts is not None else default_opts()
    n = batch.n_candidates
    res = (Result * n)()
    _check(lib().aff_predict_batch(scene.device_scene(device), batch.tasks_ptr, batch.n_tasks, batch.constraints_ptr,
                                   batch.n_constraints, batch.candidates_ptr, n, C.byref(opts), res), "aff_predict_batch")
    return list(res)


def predict_batch_multi(scene, batch, devices, opts=None):
    """aff_predict_batch_multi: one host process, candidates split statically over `devices`."""
    opts = opts if opts is not None else default_opts()
    n = batch.n_candidates
    res = (Result * n)()
    dev = (C.c_int * len(devices))(*devices)
    _check(lib().aff_predict_batch_multi(scene.desc, dev, len(devices), batch.tasks_ptr, batch.n_tasks, batch.constraints_ptr,
                                         batch.n_constraints, batch.candidates_ptr, n, C.byref(opts), res), "aff_predict_batch_multi")
    return list(res)


def rank(results):
    n = len(results)
    arr = (Result * n)(*results)
    order = (C.c_int32 * n)()
    lib().aff_rank_results(arr, n, order)
    return list(order)
