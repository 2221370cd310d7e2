"""Host-side mirror of the reference's operator interface for the AGD hot path, over the C-ABI.

Names, argument meaning and error behaviour follow
/root/reference/src/main/scala/org/apache/spark/mllib/optimization/AcceleratedGradientDescent.scala
(class AcceleratedGradientDescent :41-144, object AcceleratedGradientDescent.run :177-338) and the
spark-mllib 1.3.0 plug-in types it is given (Gradient / Updater).  The Scala/JVM facade a Spark
maintainer would compile is shipped as source under jvm/ (no JVM in this image); this module is the
executable mirror used by the tests and the benchmark.

  sc = Context(devices=[0])                      # ~ SparkContext: which GPUs, which communicator
  data = sc.parallelize(labels, X).cache()       # ~ RDD[(Double, Vector)] pinned in HBM (Suite.scala:51)
  w, loss = AcceleratedGradientDescent.run(data, LogisticGradient(), SimpleUpdater(), 1e-12, 10, 0.0,
                                           w0, 1.0, float("inf"), 0.5, 0.9, True)      # Suite.scala:62-74
  w = AcceleratedGradientDescent(LogisticGradient(), SquaredL2Updater()).setRegParam(0.2).optimize(data, w0)
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _native as N


# --------------------------------------------------------------------------- plug-in types
class Gradient:
    """org.apache.spark.mllib.optimization.Gradient [mllib-1.3.0]: closed set of built-ins."""
    kind: int = -1


class LogisticGradient(Gradient):
    kind = N.GRAD_LOGISTIC

    def __init__(self, numClasses: int = 2):
        if numClasses != 2:  # the reference only ever builds the binary form (Suite.scala:39,251)
            raise NotImplementedError("multinomial LogisticGradient is outside the reference's AGD path")


class LeastSquaresGradient(Gradient):
    """1.3.0 definition: loss diff^2, gradient 2*diff*x.  half=True selects the Spark>=1.4 definition."""

    def __init__(self, half: bool = False):
        self.kind = N.GRAD_LEAST_SQUARES_HALF if half else N.GRAD_LEAST_SQUARES


class HingeGradient(Gradient):
    kind = N.GRAD_HINGE


class Updater:
    """org.apache.spark.mllib.optimization.Updater [mllib-1.3.0]: closed set of built-ins."""
    kind: int = -1


class SimpleUpdater(Updater):
    kind = N.UPD_SIMPLE


class SquaredL2Updater(Updater):
    kind = N.UPD_SQUARED_L2


class L1Updater(Updater):
    kind = N.UPD_L1


def _grad_kind(g) -> int:
    if not isinstance(g, Gradient) or g.kind < 0:
        # the JVM facade throws UnsupportedOperationException here: there is no CPU fallback
        raise TypeError(f"unsupported Gradient {type(g).__name__}: only Logistic/LeastSquares/Hinge run on the GPU path")
    return g.kind


def _upd_kind(u) -> int:
    if not isinstance(u, Updater) or u.kind < 0:
        raise TypeError(f"unsupported Updater {type(u).__name__}: only Simple/SquaredL2/L1 run on the GPU path")
    return u.kind


_DT = {np.dtype(np.float64): N.F64, np.dtype(np.float32): N.F32}
_STORE = {"f32": N.F32, "f64": N.F64, "bf16": N.BF16, N.F32: N.F32, N.F64: N.F64, N.BF16: N.BF16}


def bf16_to_f32(raw: np.ndarray) -> np.ndarray:
    """Exact widening of raw bf16 bit patterns (uint16) to float32."""
    return (raw.astype(np.uint32) << 16).view(np.float32)


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


# --------------------------------------------------------------------------- context + dataset
class Context:
    """Which local GPUs this process drives and how its ranks sit in the world (one rank per GPU).

    devices: local CUDA ordinals.  world_size/first_rank: global layout.  Multi-process worlds
    (world_size > len(devices)) pick how the exchange is SET UP (the per-pass data never touches a library):
      transport="nccl": id_exchange(bytes|None) -> bytes returns rank 0's 128-byte NCCL id on every process
                        (e.g. a torch.distributed broadcast); NCCL ships the CUDA IPC handles and is the fallback;
      transport="ipc":  handle_exchange(bytes) -> bytes returns every process's handle blob concatenated in rank
                        order (an all-gather in the host language); no NCCL anywhere -- also works for several
                        processes sharing one GPU."""

    def __init__(self, devices: Sequence[int] = (0,), world_size: Optional[int] = None, first_rank: int = 0,
                 id_exchange=None, handle_exchange=None, transport: Optional[str] = None):
        self.devices = list(devices)
        self.world_size = len(self.devices) if world_size is None else int(world_size)
        self.first_rank = int(first_rank)
        self.id_exchange = id_exchange
        self.handle_exchange = handle_exchange
        self.transport = transport or ("ipc" if (handle_exchange is not None and id_exchange is None) else "nccl")
        if self.transport not in ("nccl", "ipc"):
            raise ValueError("transport must be 'nccl' or 'ipc'")
        if self.world_size > len(self.devices):
            if self.transport == "nccl" and id_exchange is None:
                raise ValueError("multi-process worlds need id_exchange to ship the NCCL unique id")
            if self.transport == "ipc" and handle_exchange is None:
                raise ValueError("transport='ipc' needs handle_exchange to ship the CUDA IPC handles")

    @staticmethod
    def from_torch_distributed(local_device: Optional[int] = None, transport: str = "nccl") -> "Context":
        """One process per GPU under torchrun: ranks/ids/handles travel over the existing process group."""
        import torch
        import torch.distributed as dist
        rank, world = dist.get_rank(), dist.get_world_size()
        dev = torch.cuda.current_device() if local_device is None else local_device

        def exchange(my_id):
            buf = [my_id]
            dist.broadcast_object_list(buf, src=0)
            return buf[0]

        def gather_handles(blob):
            out = [None] * world
            dist.all_gather_object(out, blob)
            return b"".join(out)

        return Context([dev], world_size=world, first_rank=rank, id_exchange=exchange, handle_exchange=gather_handles,
                       transport=transport)

    def _new_handle(self) -> C.c_void_p:
        L = N.lib()
        ids = (C.c_int32 * len(self.devices))(*self.devices)
        h = C.c_void_p()
        N.check(L.agd_create(ids, len(self.devices), C.byref(h)), None)
        if self.world_size > len(self.devices) and self.transport == "ipc":
            N.check(L.agd_comm_init_ipc(h, self.world_size, self.first_rank), h)
        elif self.world_size > len(self.devices):
            my_id = None
            if self.first_rank == 0:
                buf = C.create_string_buffer(128)
                N.check(L.agd_comm_unique_id(buf), None)
                my_id = buf.raw
            the_id = self.id_exchange(my_id)
            N.check(L.agd_comm_init(h, C.c_char_p(the_id), self.world_size, self.first_rank), h)
        return h

    # --- RDD construction (mirrors sc.parallelize(data, numSlices).cache(), Suite.scala:51) ---
    def parallelize(self, labels, X, store: str = "f64") -> "DeviceDataset":
        """Rows of THIS process are split contiguously over its local GPUs."""
        ds = DeviceDataset(self)
        ds.load_dense(labels, X, store=store)
        return ds

    def parallelize_csr(self, labels, rowptr, idx, val, d: int, store: str = "f64") -> "DeviceDataset":
        ds = DeviceDataset(self)
        ds.load_csr(labels, rowptr, idx, val, d, store=store)
        return ds

    def synthetic(self, total_rows: int, d: int, gradient: Gradient, seed: int = 42, store: str = "f32") -> "DeviceDataset":
        """The benchmark workload of SURVEY.md 8(d), generated in place on every GPU rank."""
        ds = DeviceDataset(self)
        N.check(N.lib().agd_generate(ds.h, total_rows, d, _STORE[store], seed, _grad_kind(gradient)), ds.h)
        ds.total_rows = total_rows
        return ds


def _synthetic_csr(ctx, total_rows, d, nnz_per_row, gradient, seed=42, store="f32"):
    ds = DeviceDataset(ctx)
    N.check(N.lib().agd_generate_csr(ds.h, total_rows, d, nnz_per_row, _STORE[store], seed, _grad_kind(gradient)), ds.h)
    ds.total_rows = total_rows
    return ds


Context.synthetic_csr = _synthetic_csr


class DeviceDataset:
    """RDD[(Double, Vector)] stand-in: row shards pinned in HBM for the lifetime of the object."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.h = ctx._new_handle()
        self.total_rows = 0
        self._xchg_d = 0       # dimension the host-shipped exchange (transport="ipc") was set up for

    def _ensure_exchange(self):
        """transport="ipc": ship the CUDA IPC handles of the exchange buffers once per loaded dimension (collective:
        every process reaches this from the same compute call)."""
        c = self.ctx
        if c.transport != "ipc" or c.world_size <= len(c.devices) or self._xchg_d == self.d:
            return
        L = N.lib()
        cap = len(c.devices) * N.XCHG_HANDLE_BYTES
        blob = C.create_string_buffer(cap)
        n = C.c_int64()
        N.check(L.agd_xchg_export(self.h, blob, cap, C.byref(n)), self.h)
        everyone = c.handle_exchange(blob.raw[:n.value])
        N.check(L.agd_xchg_import(self.h, C.c_char_p(everyone), len(everyone)), self.h)
        self._xchg_d = self.d

    def cache(self) -> "DeviceDataset":  # shards are always resident; kept for call-site parity
        return self

    def unpersist(self) -> "DeviceDataset":
        """Drops every shard but keeps the context (devices, communicator) for the next load."""
        N.check(N.lib().agd_clear(self.h), self.h)
        self.total_rows = 0
        self._xchg_d = 0
        return self

    def load_dense(self, labels, X, store: str = "f64"):
        labels = np.ascontiguousarray(labels, dtype=np.float64)
        X = np.asarray(X)
        if X.dtype not in _DT:
            X = X.astype(np.float64)
        if X.ndim != 2 or X.shape[0] != labels.shape[0]:
            raise ValueError("X must be (rows, d) with one label per row")
        if not X.flags.c_contiguous:
            X = np.ascontiguousarray(X)
        n, d = X.shape
        nd = len(self.ctx.devices)
        L = N.lib()
        for i in range(nd):
            lo, hi = (i * n) // nd, ((i + 1) * n) // nd
            xs, ls = X[lo:hi], labels[lo:hi]
            N.check(L.agd_load_dense(self.h, i, _ptr(xs) if hi > lo else None, _DT[X.dtype],
                                     _ptr(ls) if hi > lo else None, hi - lo, d, d, _STORE[store]), self.h)
        self.total_rows += n

    def load_csr(self, labels, rowptr, idx, val, d: int, store: str = "f64"):
        labels = np.ascontiguousarray(labels, dtype=np.float64)
        rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        val = np.ascontiguousarray(val)
        if val.dtype not in _DT:
            val = val.astype(np.float64)
        n = labels.shape[0]
        nd = len(self.ctx.devices)
        L = N.lib()
        for i in range(nd):
            lo, hi = (i * n) // nd, ((i + 1) * n) // nd
            rp = np.ascontiguousarray(rowptr[lo:hi + 1] - rowptr[lo])
            a, b = int(rowptr[lo]), int(rowptr[hi])
            N.check(L.agd_load_csr(self.h, i, _ptr(rp), _ptr(idx[a:b]), _ptr(val[a:b]), _DT[val.dtype],
                                   _ptr(labels[lo:hi]), hi - lo, d, _STORE[store]), self.h)
        self.total_rows += n

    @property
    def d(self) -> int:
        return int(N.lib().agd_dim(self.h))

    def local_rows(self, dev: int = 0) -> int:
        return int(N.lib().agd_rows(self.h, dev))

    def get_rows(self, dev: int, row0: int, rows: int, dtype=np.float32):
        """Rows as stored (dtype must match the storage: float32, float64, or uint16 for raw bf16)."""
        X = np.empty((rows, self.d), dtype=dtype)
        y = np.empty(rows, dtype=np.float64)
        N.check(N.lib().agd_get_rows(self.h, dev, row0, rows, _ptr(X), _ptr(y)), self.h)
        return X, y

    def get_labels(self, dev: int, row0: int, rows: int) -> np.ndarray:
        y = np.empty(rows, dtype=np.float64)
        N.check(N.lib().agd_get_rows(self.h, dev, row0, rows, None, _ptr(y)), self.h)
        return y

    def kernel_name(self, dev: int = 0) -> str:
        """The gradient kernel this shard dispatches to (for reports)."""
        return (N.lib().agd_kernel_name(self.h, dev) or b"").decode()

    def get_csr_rows(self, dev: int, row0: int, rows: int, nnz_capacity: int, dtype=np.float32):
        rowptr = np.empty(rows + 1, dtype=np.int64)
        idx = np.empty(nnz_capacity, dtype=np.int32)
        val = np.empty(nnz_capacity, dtype=dtype)
        y = np.empty(rows, dtype=np.float64)
        N.check(N.lib().agd_get_csr_rows(self.h, dev, row0, rows, _ptr(rowptr), _ptr(idx), _ptr(val), nnz_capacity,
                                         _ptr(y)), self.h)
        n = int(rowptr[-1])
        return rowptr, idx[:n], val[:n], y

    def set_option(self, key: str, value) -> None:
        N.check(N.lib().agd_set_option(self.h, key.encode(), str(value).encode()), self.h)

    # plug-in granularity entry points
    def smooth(self, gradient: Gradient, w):
        """applySmooth (AGD.scala:192-208) with host buffers: (loss/count, grad/count, count)."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        if w.shape[0] != self.d:
            raise ValueError("weights have the wrong dimension")
        g = np.empty(self.d, dtype=np.float64)
        loss, cnt = C.c_double(), C.c_int64()
        self._ensure_exchange()
        N.check(N.lib().agd_smooth(self.h, _grad_kind(gradient), _ptr(w), C.byref(loss), _ptr(g), C.byref(cnt)), self.h)
        return loss.value, g, cnt.value

    def smooth_pair(self, gradient: Gradient, w, w2):
        """applySmooth at w plus the loss at w2 from ONE sweep over the shards (the fused form of AGD.scala:250 + :304):
        (loss/count, grad/count, count, loss2/count)."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        w2 = np.ascontiguousarray(w2, dtype=np.float64)
        if w.shape[0] != self.d or w2.shape[0] != self.d:
            raise ValueError("weights have the wrong dimension")
        g = np.empty(self.d, dtype=np.float64)
        loss, loss2, cnt = C.c_double(), C.c_double(), C.c_int64()
        self._ensure_exchange()
        N.check(N.lib().agd_smooth_pair(self.h, _grad_kind(gradient), _ptr(w), _ptr(w2), C.byref(loss), _ptr(g),
                                        C.byref(cnt), C.byref(loss2)), self.h)
        return loss.value, g, cnt.value, loss2.value

    def smooth_two(self, gradient: Gradient, w, w2):
        """Two complete applySmooth evaluations from ONE sweep over the shards: (loss, grad, count, loss2, grad2)."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        w2 = np.ascontiguousarray(w2, dtype=np.float64)
        if w.shape[0] != self.d or w2.shape[0] != self.d:
            raise ValueError("weights have the wrong dimension")
        g, g2 = np.empty(self.d, dtype=np.float64), np.empty(self.d, dtype=np.float64)
        loss, loss2, cnt = C.c_double(), C.c_double(), C.c_int64()
        self._ensure_exchange()
        N.check(N.lib().agd_smooth_two(self.h, _grad_kind(gradient), _ptr(w), _ptr(w2), C.byref(loss), _ptr(g),
                                       C.byref(cnt), C.byref(loss2), _ptr(g2)), self.h)
        return loss.value, g, cnt.value, loss2.value, g2

    def prox(self, updater: Updater, w, g, step: float, reg: float):
        """applyProjector (AGD.scala:214-222): (regVal, newWeights)."""
        w = np.ascontiguousarray(w, dtype=np.float64)
        g = np.ascontiguousarray(g, dtype=np.float64)
        out = np.empty_like(w)
        rv = C.c_double()
        N.check(N.lib().agd_prox(self.h, _upd_kind(updater), _ptr(w), _ptr(g), step, reg, w.shape[0], _ptr(out),
                                 C.byref(rv)), self.h)
        return rv.value, out

    def close(self):
        if self.h is not None:
            N.lib().agd_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MLUtils:
    """org.apache.spark.mllib.util.MLUtils [mllib-1.3.0]: the ingest side of the path."""

    @staticmethod
    def parseLibSVMFile(path: str, numFeatures: int = -1):
        """Host-only parse: (labels, rowptr, indices, values, d) with zero-based indices."""
        L = N.lib()
        obj = C.c_void_p()
        rc = L.agd_libsvm_read(path.encode(), numFeatures, C.byref(obj))
        try:
            if rc != 0:
                raise ValueError(L.agd_libsvm_error(obj).decode())
            n, d, nnz = L.agd_libsvm_rows(obj), L.agd_libsvm_dim(obj), L.agd_libsvm_nnz(obj)

            def arr(ptr, count, ctype, dtype):
                if count == 0:
                    return np.zeros(0, dtype=dtype)
                return np.ctypeslib.as_array(C.cast(ptr, C.POINTER(ctype)), shape=(count,)).astype(dtype, copy=True)

            return (arr(L.agd_libsvm_labels(obj), n, C.c_double, np.float64),
                    arr(L.agd_libsvm_rowptr(obj), n + 1, C.c_int64, np.int64),
                    arr(L.agd_libsvm_indices(obj), nnz, C.c_int32, np.int32),
                    arr(L.agd_libsvm_values(obj), nnz, C.c_double, np.float64), int(d))
        finally:
            L.agd_libsvm_free(obj)

    @staticmethod
    def loadLibSVMFile(sc: "Context", path: str, numFeatures: int = -1, store: str = "f64") -> "DeviceDataset":
        """loadLibSVMFile(sc, path, numFeatures): rows land as CSR shards on the context's GPUs."""
        ds = DeviceDataset(sc)
        N.check(N.lib().agd_load_libsvm(ds.h, path.encode(), numFeatures, _STORE[store]), ds.h)
        ds.total_rows = sum(ds.local_rows(i) for i in range(len(sc.devices)))
        return ds


@dataclass
class RunStats:
    iterations: int
    passes: int
    backtracks: int
    restarts: int
    converged: bool
    stopped_nan: bool
    nonterminating: bool
    final_L: float
    final_theta: float
    seconds_total: float
    k1_ms_total: float
    k1_launches: int
    gpu_launches: int
    allreduce_ms_total: float
    device_ms_total: float
    collective_calls: int
    collective_kind: int = 0
    fused_passes: int = 0      # evaluations that shared a sweep over X with another one (sweeps = passes - fused_passes)
    wasted_passes: int = 0


def _stats(st: N.Stats) -> RunStats:
    return RunStats(st.iterations, st.passes, st.backtracks, st.restarts, bool(st.converged), bool(st.stopped_nan),
                    bool(st.nonterminating), st.final_L, st.final_theta, st.seconds_total, st.k1_ms_total,
                    st.k1_launches, st.gpu_launches, st.allreduce_ms_total, st.device_ms_total, st.collective_calls, st.collective_kind,
                    st.fused_passes, st.wasted_passes)


# --------------------------------------------------------------------------- the optimizer
class AcceleratedGradientDescent:
    """class AcceleratedGradientDescent(gradient, updater) extends Optimizer (AGD.scala:41-144)."""

    def __init__(self, gradient: Gradient, updater: Updater):
        self.gradient = gradient
        self.updater = updater
        self.convergenceTol = 1e-4          # AGD.scala:44
        self.numIterations = 100            # :45
        self.regParam = 0.0                 # :46
        self.L0 = 1.0                       # :47
        self.Lexact = float("inf")          # :48
        self.beta = 0.5                     # :49
        self.alpha = 0.9                    # :50
        self.mayRestart = True              # :51
        self.memoize = False                # extension: AGD_FLAG_MEMOIZE_FX (bit-identical, fewer passes)
        self.fuse = True                    # pass fusion (AGD_FLAG_NO_FUSE clears it): same evaluations, fewer sweeps
        self.last_stats: Optional[RunStats] = None

    def setConvergenceTol(self, tol: float): self.convergenceTol = tol; return self       # :57
    def setNumIterations(self, iters: int): self.numIterations = iters; return self       # :65
    def setRegParam(self, regParam: float): self.regParam = regParam; return self         # :73
    def setL0(self, L0: float): self.L0 = L0; return self                                 # :78
    def setLexact(self, Lexact: float): self.Lexact = Lexact; return self                 # :83
    def setBeta(self, beta: float): self.beta = beta; return self                         # :88
    def setAlpha(self, alpha: float): self.alpha = alpha; return self                     # :93
    def setMayRestart(self, mayRestart: bool): self.mayRestart = mayRestart; return self  # :98
    def setGradient(self, gradient: Gradient): self.gradient = gradient; return self      # :106
    def setUpdater(self, updater: Updater): self.updater = updater; return self           # :117
    def setMemoize(self, on: bool): self.memoize = on; return self
    def setFuse(self, on: bool): self.fuse = on; return self

    def optimize(self, data: DeviceDataset, initialWeights) -> np.ndarray:                # :128-143
        w, _, st = run_with_stats(data, self.gradient, self.updater, self.convergenceTol, self.numIterations,
                                  self.regParam, initialWeights, self.L0, self.Lexact, self.beta, self.alpha,
                                  self.mayRestart, memoize=self.memoize, fuse=self.fuse)
        self.last_stats = st
        return w

    @staticmethod
    def run(data: DeviceDataset, gradient: Gradient, updater: Updater, convergenceTol: float, numIterations: int,
            regParam: float, initialWeights, L0: float, Lexact: float, beta: float, alpha: float,
            mayRestart: bool):
        """object AcceleratedGradientDescent.run (AGD.scala:177-189): returns (weights, lossHistory)."""
        w, hist, _ = run_with_stats(data, gradient, updater, convergenceTol, numIterations, regParam,
                                    initialWeights, L0, Lexact, beta, alpha, mayRestart)
        return w, hist


def run_with_stats(data: DeviceDataset, gradient, updater, convergenceTol, numIterations, regParam, initialWeights,
                   L0=1.0, Lexact=float("inf"), beta=0.5, alpha=0.9, mayRestart=True, memoize=False, fuse=True):
    if not isinstance(data, DeviceDataset):
        raise TypeError("data must be a DeviceDataset (Context.parallelize(...)); there is no CPU path")
    w0 = np.ascontiguousarray(initialWeights, dtype=np.float64)
    if w0.ndim != 1 or w0.shape[0] != data.d:
        raise ValueError(f"initialWeights has size {w0.shape}, data has {data.d} features")
    p = N.Params(convergenceTol, int(numIterations), regParam, L0, Lexact, beta, alpha, int(bool(mayRestart)),
                 _grad_kind(gradient), _upd_kind(updater), (N.FLAG_MEMOIZE_FX if memoize else 0) | (0 if fuse else N.FLAG_NO_FUSE))
    w = np.empty_like(w0)
    hist = np.empty(max(int(numIterations), 1), dtype=np.float64)
    nh, st = C.c_int32(), N.Stats()
    data._ensure_exchange()
    N.check(N.lib().agd_run(data.h, C.byref(p), _ptr(w0), _ptr(w), _ptr(hist), C.byref(nh), C.byref(st)), data.h)
    return w, hist[:nh.value].copy(), _stats(st)


class GradientDescent:
    """GradientDescent.runMiniBatchSGD [mllib-1.3.0], the comparator of Suite.scala:78,118,225
    (miniBatchFraction < 1 samples rows with a counter-based Bernoulli mask keyed by 42 + i, see include/agd_b200.h)."""

    @staticmethod
    def runMiniBatchSGD(data: DeviceDataset, gradient: Gradient, updater: Updater, stepSize: float, numIterations: int,
                        regParam: float, miniBatchFraction: float, initialWeights):
        if not isinstance(data, DeviceDataset):
            raise TypeError("data must be a DeviceDataset (Context.parallelize(...)); there is no CPU path")
        w0 = np.ascontiguousarray(initialWeights, dtype=np.float64)
        if w0.ndim != 1 or w0.shape[0] != data.d:   # the native side reads and writes agd_dim(h) doubles
            raise ValueError(f"initialWeights has size {w0.shape}, data has {data.d} features")
        w = np.empty_like(w0)
        hist = np.empty(max(int(numIterations), 1), dtype=np.float64)
        nh, st = C.c_int32(), N.Stats()
        data._ensure_exchange()
        N.check(N.lib().agd_gd_run_minibatch(data.h, _grad_kind(gradient), _upd_kind(updater), stepSize,
                                             int(numIterations), regParam, float(miniBatchFraction), _ptr(w0), _ptr(w),
                                             _ptr(hist), C.byref(nh), C.byref(st)), data.h)
        return w, hist[:nh.value].copy()
