"""ctypes binding of the CPU oracle (oracle/liboracle.so).  Test infrastructure:
only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs import
this; nothing in the product does."""
import ctypes as C
import os

import numpy as np

from affaction_b200 import Candidate, Opts, Result, TaskDesc, REPO_ROOT

ORACLE_DIR = os.path.join(REPO_ROOT, "oracle")

_libs = {}


def oracle_lib(libm=False):
    name = "liboracle_libm.so" if libm else "liboracle.so"
    if name not in _libs:
        path = os.path.join(ORACLE_DIR, name)
        if not os.path.isfile(path):
            raise RuntimeError("%s missing: run `make oracle`" % path)
        L = C.CDLL(path)
        vp, sz = C.c_void_p, C.c_size_t
        L.oracle_predict.restype = C.c_int
        L.oracle_predict.argtypes = [vp, vp, vp, vp, vp, vp, vp, sz]
        L.oracle_predict_batch.restype = C.c_int
        L.oracle_predict_batch.argtypes = [vp, vp, vp, vp, sz, vp, C.c_int, vp]
        L.oracle_num_steps.restype = C.c_int
        L.oracle_num_steps.argtypes = [vp, vp, vp]
        L.oracle_fk.restype = C.c_int
        L.oracle_fk.argtypes = [vp, vp, vp]
        L.oracle_task_eval.restype = C.c_int
        L.oracle_task_eval.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_task_dx.restype = C.c_int
        L.oracle_task_dx.argtypes = [vp, vp, vp, vp, vp]
        L.oracle_shape_distance.restype = C.c_double
        L.oracle_shape_distance.argtypes = [C.c_int, vp, vp, C.c_int, vp, vp]
        L.oracle_via_step.restype = None
        L.oracle_via_step.argtypes = [vp, vp, C.c_int, C.c_double, C.c_double]
        L.oracle_right_inverse.restype = C.c_int
        L.oracle_right_inverse.argtypes = [C.c_int, C.c_int, vp, vp, vp, vp, C.c_double, vp]
        L.oracle_uses_libm.restype = C.c_int
        _libs[name] = L
    return _libs[name]


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def predict(scene, batch, cand, opts, log_q=False, libm=False):
    """One candidate through the oracle -> (Result, q_log or None)."""
    L = oracle_lib(libm)
    r = Result()
    q = None
    rows = 0
    if log_q:
        rows = L.oracle_num_steps(batch.constraints_ptr, batch.candidate_ptr(cand), C.byref(opts)) + 1
        q = np.zeros((rows, scene.nJ))
    rc = L.oracle_predict(scene.desc, batch.tasks_ptr, batch.constraints_ptr, batch.candidate_ptr(cand), C.byref(opts),
                          C.byref(r), _p(q) if q is not None else None, rows)
    if rc != 0:
        raise RuntimeError("oracle_predict failed")
    r.idx = cand
    return r, q


def predict_batch(scene, batch, opts, n_threads=1, libm=False):
    L = oracle_lib(libm)
    n = batch.n_candidates
    res = (Result * n)()
    rc = L.oracle_predict_batch(scene.desc, batch.tasks_ptr, batch.constraints_ptr, batch.candidates_ptr, n, C.byref(opts),
                                n_threads, res)
    if rc != 0:
        raise RuntimeError("oracle_predict_batch failed")
    return list(res)


def fk(scene, batch=None, cand=None):
    A = np.zeros((scene.n_bodies, 12))
    cp = batch.candidate_ptr(cand) if batch is not None else None
    oracle_lib().oracle_fk(scene.desc, cp, _p(A))
    return A


def task_eval(scene, task, q=None):
    x = np.zeros(8)
    J = np.zeros((8, scene.nJ))
    dim = oracle_lib().oracle_task_eval(scene.desc, C.byref(task), _p(q) if q is not None else None, _p(x), _p(J))
    return x[:dim].copy(), J[:dim].copy()


def task_dx(scene, task, q, x_des):
    dx = np.zeros(8)
    xd = np.ascontiguousarray(x_des, dtype=np.float64)
    dim = oracle_lib().oracle_task_dx(scene.desc, C.byref(task), _p(q) if q is not None else None, _p(xd), _p(dx))
    return dx[:dim].copy()


def shape_distance(t1, ext1, A1, t2, ext2, A2):
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (ext1, A1, ext2, A2)]
    return oracle_lib().oracle_shape_distance(t1, _p(a[0]), _p(a[1]), t2, _p(a[2]), _p(a[3]))


def via_step(state, vias, horizon=1.0, dt=0.01):
    st = np.ascontiguousarray(state, dtype=np.float64).copy()
    v = np.ascontiguousarray(vias, dtype=np.float64).reshape(-1, 5)
    oracle_lib().oracle_via_step(_p(st), _p(v), len(v), horizon, dt)
    return st


def right_inverse(J, w, dx, dH, lam):
    J = np.ascontiguousarray(J, dtype=np.float64)
    m, n = J.shape
    dq = np.zeros(n)
    ok = oracle_lib().oracle_right_inverse(m, n, _p(J), _p(np.ascontiguousarray(w, dtype=np.float64)),
                                           _p(np.ascontiguousarray(dx, dtype=np.float64)),
                                           _p(np.ascontiguousarray(dH, dtype=np.float64)), lam, _p(dq))
    return ok, dq
