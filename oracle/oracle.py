"""ctypes binding of the CPU oracle (oracle/agd_oracle.c, oracle/synth_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")

GRAD = {"logistic": 0, "least_squares": 1, "hinge": 2, "least_squares_half": 3}
UPD = {"simple": 0, "squared_l2": 1, "l1": 2}
DENSE_F64, DENSE_F32, CSR = 0, 1, 2


def build(force: bool = False) -> str:
    """Compile liboracle.so with the committed Makefile (gcc, -ffp-contract=off, OpenMP)."""
    srcs = [os.path.join(_HERE, f) for f in ("agd_oracle.c", "synth_oracle.c", "agd_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


class _Data(C.Structure):
    _fields_ = [("n", C.c_int64), ("d", C.c_int32), ("storage", C.c_int32),
                ("Xd", C.c_void_p), ("Xf", C.c_void_p), ("ld", C.c_int64),
                ("rowptr", C.c_void_p), ("csr_idx", C.c_void_p), ("csr_val", C.c_void_p),
                ("csr_val_f32", C.c_void_p), ("labels", C.c_void_p), ("sample_seed", C.c_uint64),
                ("sample_thresh", C.c_uint64)]


class _Params(C.Structure):
    _fields_ = [("convergence_tol", C.c_double), ("num_iterations", C.c_int32), ("reg_param", C.c_double),
                ("L0", C.c_double), ("Lexact", C.c_double), ("beta", C.c_double), ("alpha", C.c_double),
                ("may_restart", C.c_int32), ("partitions", C.c_int32), ("threads", C.c_int32)]


class _Stats(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("passes", C.c_int32), ("backtracks", C.c_int32),
                ("restarts", C.c_int32), ("converged", C.c_int32), ("stopped_nan", C.c_int32),
                ("nonterminating", C.c_int32), ("final_L", C.c_double), ("final_theta", C.c_double)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_smooth.restype = C.c_int
        _lib.oracle_agd_run.restype = C.c_int
        _lib.oracle_gd_run.restype = C.c_int
        _lib.oracle_max_threads.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Data:
    """Host-resident RDD[(Double, Vector)] stand-in.  Keeps numpy arrays alive."""

    def __init__(self, labels, X=None, csr=None, d=None):
        self.labels = np.ascontiguousarray(labels, dtype=np.float64)
        self.n = int(self.labels.shape[0])
        s = _Data()
        s.n = self.n
        s.labels = _p(self.labels)
        if X is not None:
            assert X.ndim == 2 and X.shape[0] == self.n
            if X.dtype == np.float32:
                self.X = np.ascontiguousarray(X)
                s.storage, s.Xf = DENSE_F32, _p(self.X)
            else:
                self.X = np.ascontiguousarray(X, dtype=np.float64)
                s.storage, s.Xd = DENSE_F64, _p(self.X)
            s.d, s.ld = int(X.shape[1]), int(X.shape[1])
        else:
            rowptr, idx, val = csr
            self.rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
            self.idx = np.ascontiguousarray(idx, dtype=np.int32)
            s.storage, s.d = CSR, int(d)
            s.rowptr, s.csr_idx = _p(self.rowptr), _p(self.idx)
            if val.dtype == np.float32:
                self.val = np.ascontiguousarray(val)
                s.csr_val_f32 = _p(self.val)
            else:
                self.val = np.ascontiguousarray(val, dtype=np.float64)
                s.csr_val = _p(self.val)
        self.d = int(s.d)
        self._s = s


@dataclass
class RunResult:
    weights: np.ndarray
    loss_history: np.ndarray
    iterations: int
    passes: int
    backtracks: int
    restarts: int
    converged: bool
    stopped_nan: bool
    nonterminating: bool
    final_L: float
    final_theta: float


def smooth(data: Data, gradient: str, w, partitions: int = 2, threads: int = 1):
    """applySmooth (AGD.scala:192-208): returns (loss/count, grad/count, count)."""
    w = np.ascontiguousarray(w, dtype=np.float64)
    g = np.empty(data.d, dtype=np.float64)
    loss, cnt = C.c_double(), C.c_int64()
    rc = lib().oracle_smooth(C.byref(data._s), GRAD[gradient], _p(w), partitions, threads,
                             C.byref(loss), _p(g), C.byref(cnt))
    assert rc == 0
    return loss.value, g, cnt.value


def prox(updater: str, w, g, step: float, reg: float):
    """applyProjector (AGD.scala:214-222): returns (regVal, newWeights)."""
    w = np.ascontiguousarray(w, dtype=np.float64)
    g = np.ascontiguousarray(g, dtype=np.float64)
    out = np.empty_like(w)
    rv = C.c_double()
    lib().oracle_prox(UPD[updater], _p(w), _p(g), C.c_double(step), C.c_double(reg), C.c_int32(w.size),
                      _p(out), C.byref(rv))
    return rv.value, out


def agd_run(data: Data, gradient: str, updater: str, w0, *, convergence_tol=1e-4, num_iterations=100,
            reg_param=0.0, L0=1.0, Lexact=float("inf"), beta=0.5, alpha=0.9, may_restart=True,
            partitions=2, threads=1) -> RunResult:
    """AcceleratedGradientDescent.run (AGD.scala:177-338)."""
    p = _Params(convergence_tol, num_iterations, reg_param, L0, Lexact, beta, alpha, int(may_restart),
                partitions, threads)
    w0 = np.ascontiguousarray(w0, dtype=np.float64)
    w = np.empty_like(w0)
    hist = np.empty(max(num_iterations, 1), dtype=np.float64)
    nh, st = C.c_int32(), _Stats()
    rc = lib().oracle_agd_run(C.byref(data._s), GRAD[gradient], UPD[updater], C.byref(p), _p(w0), _p(w),
                              _p(hist), C.byref(nh), C.byref(st))
    assert rc == 0
    return RunResult(w, hist[:nh.value].copy(), st.iterations, st.passes, st.backtracks, st.restarts,
                     bool(st.converged), bool(st.stopped_nan), bool(st.nonterminating), st.final_L,
                     st.final_theta)


def gd_run(data: Data, gradient: str, updater: str, w0, *, step_size=1.0, num_iterations=100, reg_param=0.0,
           partitions=2, threads=1, mini_batch_fraction=1.0):
    """GradientDescent.runMiniBatchSGD (comparator of Suite.scala:78); fraction < 1 uses the counter-based row mask."""
    w0 = np.ascontiguousarray(w0, dtype=np.float64)
    w = np.empty_like(w0)
    hist = np.empty(max(num_iterations, 1), dtype=np.float64)
    nh = C.c_int32()
    if mini_batch_fraction < 1.0:
        lib().oracle_gd_run_minibatch.restype = C.c_int
        rc = lib().oracle_gd_run_minibatch(C.byref(data._s), GRAD[gradient], UPD[updater], C.c_double(step_size),
                                           C.c_int(num_iterations), C.c_double(reg_param),
                                           C.c_double(mini_batch_fraction), C.c_int(partitions), C.c_int(threads),
                                           _p(w0), _p(w), _p(hist), C.byref(nh))
        assert rc == 0
        return w, hist[:nh.value].copy()
    rc = lib().oracle_gd_run(C.byref(data._s), GRAD[gradient], UPD[updater], C.c_double(step_size),
                             C.c_int(num_iterations), C.c_double(reg_param), C.c_int(partitions),
                             C.c_int(threads), _p(w0), _p(w), _p(hist), C.byref(nh))
    assert rc == 0
    return w, hist[:nh.value].copy()


def generate_gd_input(offset: float, scale: float, n_points: int, seed: int):
    """GradientDescentSuite.generateGDInput (called at Suite.scala:46): returns (x1, y)."""
    x1 = np.empty(n_points, dtype=np.float64)
    y = np.empty(n_points, dtype=np.float64)
    lib().oracle_generate_gd_input(C.c_double(offset), C.c_double(scale), C.c_int32(n_points),
                                   C.c_int32(seed), _p(x1), _p(y))
    return x1, y


def jrandom_doubles(seed: int, n: int) -> np.ndarray:
    """n successive java.util.Random(seed).nextDouble() draws."""
    out = np.empty(n, dtype=np.float64)
    lib().oracle_jrandom_fill_double(C.c_int64(seed), C.c_int64(n), _p(out))
    return out


def synth_dense_f32(seed: int, row0: int, rows: int, d: int) -> np.ndarray:
    X = np.empty((rows, d), dtype=np.float32)
    lib().oracle_synth_dense_f32(C.c_uint64(seed), C.c_int64(row0), C.c_int64(rows), C.c_int32(d), _p(X))
    return X


def synth_wtrue(seed: int, d: int) -> np.ndarray:
    w = np.empty(d, dtype=np.float64)
    lib().oracle_synth_wtrue(C.c_uint64(seed), C.c_int32(d), _p(w))
    return w


def synth_labels(seed: int, gradient: str, row0: int, X: np.ndarray, w_true: np.ndarray) -> np.ndarray:
    y = np.empty(X.shape[0], dtype=np.float64)
    lib().oracle_synth_labels(C.c_uint64(seed), C.c_int(GRAD[gradient]), C.c_int64(row0), C.c_int64(X.shape[0]),
                              C.c_int32(X.shape[1]), _p(X), _p(np.ascontiguousarray(w_true, dtype=np.float64)), _p(y))
    return y


def synth_csr_f32(seed: int, row0: int, rows: int, d: int, k: int):
    rowptr = np.empty(rows + 1, dtype=np.int64)
    idx = np.empty(rows * k, dtype=np.int32)
    val = np.empty(rows * k, dtype=np.float32)
    lib().oracle_synth_csr_f32(C.c_uint64(seed), C.c_int64(row0), C.c_int64(rows), C.c_int32(d), C.c_int32(k),
                               _p(rowptr), _p(idx), _p(val))
    return rowptr, idx, val


def max_threads() -> int:
    return lib().oracle_max_threads()


def host_threads() -> int:
    """CPUs this process may run on (not OMP_NUM_THREADS: torchrun exports OMP_NUM_THREADS=1 to every rank)."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:
        return max(1, os.cpu_count() or 1)


def bind_threads(threads: int) -> int:
    """Pins OpenMP thread t to the t-th CPU of the affinity mask (CPU-timing hygiene; see agd_oracle.c)."""
    lib().oracle_set_threads(C.c_int(threads))
    lib().oracle_bind_threads.restype = C.c_int
    return lib().oracle_bind_threads(C.c_int(threads))


def unbind_threads() -> None:
    lib().oracle_unbind_threads()


def first_touch(a: np.ndarray, partitions: int, threads: int) -> np.ndarray:
    """Zero-fills a fresh (rows, ...) array partition by partition in oracle_smooth's thread mapping."""
    rows = a.shape[0]
    lib().oracle_first_touch(_p(a), C.c_int64(rows), C.c_int64(a.nbytes // max(rows, 1)), C.c_int(partitions),
                             C.c_int(threads))
    return a


def synth_dense_f32_placed(seed: int, row0: int, rows: int, d: int, partitions: int, threads: int) -> np.ndarray:
    """synth_dense_f32 written partition by partition by the threads that will fold those partitions."""
    X = np.empty((rows, d), dtype=np.float32)
    lib().oracle_synth_dense_f32_placed(C.c_uint64(seed), C.c_int64(row0), C.c_int64(rows), C.c_int32(d), _p(X),
                                        C.c_int(partitions), C.c_int(threads))
    return X
