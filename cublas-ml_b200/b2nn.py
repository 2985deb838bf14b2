"""ctypes binding of the C ABI in include/b2nn.h — what tests/ and bench.py call.

This is the thinnest possible host layer: it does no arithmetic and has no fallback.  If libb2nn.so is missing, or
no sm_100 device is present, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb2nn.so")

ACT_SIGMOID, ACT_TANH, ACT_CUSTOM = 0, 1, 2
OUT_SIGMOID, OUT_SOFTMAX, OUT_LINEAR, OUT_CUSTOM = 0, 1, 2, 3
COST_NEGLNMAX, COST_CROSSENTROPY, COST_SQUARED, COST_CUSTOM = 0, 1, 2, 3
INIT_1, INIT_2, INIT_CUSTOM = 1, 2, 3
PREC_FP32, PREC_TF32, PREC_AUTO, PREC_TF32X3, PREC_BF16 = 0, 1, 2, 3, 4
EPI_STORE, EPI_SIGMOID, EPI_TANH, EPI_DSIGMOID, EPI_DTANH = 0, 1, 2, 3, 4

ACT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int)
OUT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_int, C.c_int)
COST_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int)
INIT_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_int, C.c_int)
LOG_CB = C.CFUNCTYPE(None, C.c_void_p, C.c_int, C.c_float)


class TrainDesc(C.Structure):
    _fields_ = [("momentum", C.c_float), ("rate", C.c_float), ("iters", C.c_int), ("display", C.c_int),
                ("batch_num", C.c_int), ("classify", C.c_int), ("act_kind", C.c_int), ("out_kind", C.c_int),
                ("cost_kind", C.c_int), ("precision", C.c_int), ("skip_final_cost", C.c_int), ("reserved0", C.c_int),
                ("act_fn", C.c_void_p), ("act_grad_fn", C.c_void_p), ("out_fn", C.c_void_p), ("cost_fn", C.c_void_p)]


class TrainResult(C.Structure):
    _fields_ = [("call_seconds", C.c_double), ("loop_seconds", C.c_double), ("final_cost", C.c_double),
                ("error_count", C.c_double), ("steps", C.c_int64), ("kernel_launches", C.c_int64),
                ("gemm_tc_launches", C.c_int64)]


# every symbol include/b2nn.h declares (tests check the library exports all of them)
ABI_SYMBOLS = [
    "b2nn_create", "b2nn_destroy", "b2nn_last_error", "b2nn_device_ok", "b2nn_version", "b2nn_set_layers",
    "b2nn_num_weights", "b2nn_get_weights", "b2nn_set_weights", "b2nn_set_train_data", "b2nn_train",
    "b2nn_train_step_host", "b2nn_train_host", "b2nn_reset_velocity", "b2nn_evaluate", "b2nn_predict", "b2nn_predict_device",
    "b2nn_builtin_kind", "b2nn_comm_unique_id", "b2nn_comm_init", "b2nn_comm_barrier", "b2nn_comm_p2p_export", "b2nn_comm_p2p_open", "b2nn_gemm",
    "b2nn_momentum_update", "b2nn_stream", "b2nn_profile_enable", "b2nn_profile_read",
    "b2nn_save_checkpoint", "b2nn_load_checkpoint", "b2nn_set_train_data_raw", "b2nn_set_normalisation",
    "b2nn_get_normalisation", "b2nn_evaluate_raw", "b2nn_predict_raw",
]
PROFILE_KINDS = ("forward", "backprop", "wgrad", "output", "update")

_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python cublas-ml_b200/build.py` (or __graft_entry__.build()). "
                           "There is no CPU fallback.")
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    vp, fp, i64 = C.c_void_p, C.POINTER(C.c_float), C.c_int64
    L.b2nn_create.restype = C.c_int
    L.b2nn_create.argtypes = [C.POINTER(vp), C.c_int]
    L.b2nn_destroy.restype = None
    L.b2nn_destroy.argtypes = [vp]
    L.b2nn_last_error.restype = C.c_char_p
    L.b2nn_device_ok.restype = C.c_int
    L.b2nn_version.restype = C.c_char_p
    L.b2nn_set_layers.restype = C.c_int
    L.b2nn_set_layers.argtypes = [vp, C.POINTER(C.c_int), C.c_int, C.c_int, vp, C.c_uint64, C.c_int]
    L.b2nn_num_weights.restype = i64
    L.b2nn_num_weights.argtypes = [vp]
    L.b2nn_get_weights.restype = C.c_int
    L.b2nn_get_weights.argtypes = [vp, fp, i64]
    L.b2nn_set_weights.restype = C.c_int
    L.b2nn_set_weights.argtypes = [vp, fp, i64]
    L.b2nn_set_train_data.restype = C.c_int
    L.b2nn_set_train_data.argtypes = [vp, vp, vp, i64, C.c_int, C.c_int]
    L.b2nn_train.restype = C.c_int
    L.b2nn_train.argtypes = [vp, C.POINTER(TrainDesc), vp, vp, C.POINTER(TrainResult)]
    L.b2nn_train_step_host.restype = C.c_int
    L.b2nn_train_step_host.argtypes = [vp, C.POINTER(TrainDesc), vp, vp, i64, i64, C.POINTER(C.c_float)]
    L.b2nn_train_host.restype = C.c_int
    L.b2nn_train_host.argtypes = [vp, C.POINTER(TrainDesc), vp, vp, i64, C.c_int, C.c_int, C.POINTER(C.c_float),
                                  C.POINTER(TrainResult)]
    L.b2nn_reset_velocity.restype = C.c_int
    L.b2nn_reset_velocity.argtypes = [vp]
    L.b2nn_evaluate.restype = C.c_int
    L.b2nn_evaluate.argtypes = [vp, C.POINTER(TrainDesc), vp, vp, i64, C.c_int, C.POINTER(C.c_double),
                                C.POINTER(C.c_double)]
    L.b2nn_predict.restype = C.c_int
    L.b2nn_predict.argtypes = [vp, C.POINTER(TrainDesc), vp, i64, vp, vp]
    L.b2nn_predict_device.restype = C.c_int
    L.b2nn_predict_device.argtypes = [vp, C.POINTER(TrainDesc), vp, i64, i64, vp]
    L.b2nn_builtin_kind.restype = C.c_int
    L.b2nn_builtin_kind.argtypes = [vp, C.c_int]
    L.b2nn_comm_unique_id.restype = C.c_int
    L.b2nn_comm_unique_id.argtypes = [vp]
    L.b2nn_comm_init.restype = C.c_int
    L.b2nn_comm_init.argtypes = [vp, C.c_int, C.c_int, vp]
    L.b2nn_comm_barrier.restype = C.c_int
    L.b2nn_comm_barrier.argtypes = [vp]
    L.b2nn_comm_p2p_export.restype = C.c_int
    L.b2nn_comm_p2p_export.argtypes = [vp, vp]
    L.b2nn_comm_p2p_open.restype = C.c_int
    L.b2nn_comm_p2p_open.argtypes = [vp, vp]
    L.b2nn_gemm.restype = C.c_int
    L.b2nn_gemm.argtypes = [vp, C.c_int, vp, i64, C.c_int, vp, i64, C.c_int, vp, i64, C.c_int, C.c_int, C.c_int,
                            C.c_int, vp, i64]
    L.b2nn_momentum_update.restype = C.c_int
    L.b2nn_momentum_update.argtypes = [vp, vp, vp, vp, i64, C.c_float, C.c_float]
    L.b2nn_stream.restype = vp
    L.b2nn_stream.argtypes = [vp]
    L.b2nn_profile_enable.restype = C.c_int
    L.b2nn_profile_enable.argtypes = [vp, C.c_int]
    L.b2nn_profile_read.restype = C.c_int
    L.b2nn_profile_read.argtypes = [vp, C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                    C.POINTER(i64), C.POINTER(i64)]
    L.b2nn_save_checkpoint.restype = C.c_int
    L.b2nn_save_checkpoint.argtypes = [vp, C.c_char_p]
    L.b2nn_load_checkpoint.restype = C.c_int
    L.b2nn_load_checkpoint.argtypes = [vp, C.c_char_p]
    L.b2nn_set_train_data_raw.restype = C.c_int
    L.b2nn_set_train_data_raw.argtypes = [vp, vp, vp, i64, C.c_int, C.c_int, C.c_int, vp, vp]
    L.b2nn_set_normalisation.restype = C.c_int
    L.b2nn_set_normalisation.argtypes = [vp, vp, vp, C.c_int]
    L.b2nn_get_normalisation.restype = C.c_int
    L.b2nn_get_normalisation.argtypes = [vp, vp, vp, C.c_int]
    L.b2nn_evaluate_raw.restype = C.c_int
    L.b2nn_evaluate_raw.argtypes = [vp, C.POINTER(TrainDesc), vp, vp, i64, C.c_int, C.c_int, C.POINTER(C.c_double),
                                    C.POINTER(C.c_double)]
    L.b2nn_predict_raw.restype = C.c_int
    L.b2nn_predict_raw.argtypes = [vp, C.POINTER(TrainDesc), vp, i64, C.c_int, vp, vp]
    _lib = L
    return L


class B2nnError(RuntimeError):
    pass


def _check(rc: int, what: str) -> None:
    if rc != 0:
        raise B2nnError(f"{what} failed (rc={rc}): {lib().b2nn_last_error().decode(errors='replace')}")


def device_ok() -> bool:
    return bool(lib().b2nn_device_ok())


def _host_ptr(a: np.ndarray):
    return C.c_void_p(a.ctypes.data)


@dataclass
class TrainOut:
    log_iter: np.ndarray
    log_J: np.ndarray
    final_cost: float
    error_count: float
    call_seconds: float
    loop_seconds: float
    steps: int
    kernel_launches: int
    gemm_tc_launches: int


def make_desc(*, momentum=0.9, rate=0.01, iters=1, display=1, batch_num=1, classify=True, act=ACT_TANH,
              out=OUT_SOFTMAX, cost=COST_CROSSENTROPY, precision=PREC_AUTO, skip_final_cost=False) -> TrainDesc:
    d = TrainDesc()
    d.momentum, d.rate, d.iters, d.display, d.batch_num = momentum, rate, iters, display, batch_num
    d.classify, d.act_kind, d.out_kind, d.cost_kind = int(classify), act, out, cost
    d.precision, d.skip_final_cost = precision, int(skip_final_cost)
    return d


class Net:
    """One engine context (class cublasNN's device state)."""

    def __init__(self, device: int = -1):
        self._h = C.c_void_p()
        _check(lib().b2nn_create(C.byref(self._h), device), "b2nn_create")
        self.layers: list[int] = []
        self._keep = []

    def close(self):
        if self._h:
            lib().b2nn_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def set_layers(self, layers, init_kind=INIT_2, seed=42, seed_is_clock=False, custom_init=None):
        arr = (C.c_int * len(layers))(*layers)
        fn = None
        if custom_init is not None:
            fn = C.cast(custom_init, C.c_void_p)
            self._keep.append(custom_init)
        _check(lib().b2nn_set_layers(self._h, arr, len(layers), init_kind, fn, seed, int(seed_is_clock)),
               "b2nn_set_layers")
        self.layers = list(layers)

    def num_weights(self) -> int:
        return int(lib().b2nn_num_weights(self._h))

    def get_weights(self) -> np.ndarray:
        out = np.empty(self.num_weights(), dtype=np.float32)
        _check(lib().b2nn_get_weights(self._h, out.ctypes.data_as(C.POINTER(C.c_float)), out.size), "b2nn_get_weights")
        return out

    def set_weights(self, theta: np.ndarray):
        theta = np.ascontiguousarray(theta, dtype=np.float32)
        _check(lib().b2nn_set_weights(self._h, theta.ctypes.data_as(C.POINTER(C.c_float)), theta.size),
               "b2nn_set_weights")

    def set_train_data(self, x_cm: np.ndarray, y_cm: np.ndarray, m: int):
        """x_cm: column-major m x (l0+1) with leading ones column; y_cm: column-major m x l_last (flat float32)."""
        x_cm = np.ascontiguousarray(x_cm, dtype=np.float32)
        y_cm = np.ascontiguousarray(y_cm, dtype=np.float32)
        _check(lib().b2nn_set_train_data(self._h, _host_ptr(x_cm), _host_ptr(y_cm), m, self.layers[0] + 1,
                                         self.layers[-1]), "b2nn_set_train_data")

    def set_train_data_raw(self, x_raw_cm: np.ndarray, y_cm: np.ndarray, m: int, normalise: bool = True):
        """x_raw_cm: column-major m x l0 WITHOUT the ones column; normalisation + bias happen on the device.
        Returns (mean, stddev) when normalise."""
        x_raw_cm = np.ascontiguousarray(x_raw_cm, dtype=np.float32)
        y_cm = np.ascontiguousarray(y_cm, dtype=np.float32)
        n = self.layers[0]
        mean, sd = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        _check(lib().b2nn_set_train_data_raw(self._h, _host_ptr(x_raw_cm), _host_ptr(y_cm), m, n, self.layers[-1],
                                             int(normalise), _host_ptr(mean), _host_ptr(sd)), "b2nn_set_train_data_raw")
        return (mean, sd) if normalise else None

    def set_normalisation(self, mean, sd):
        if mean is None:
            _check(lib().b2nn_set_normalisation(self._h, None, None, 0), "b2nn_set_normalisation")
            return
        mean = np.ascontiguousarray(mean, dtype=np.float32)
        sd = np.ascontiguousarray(sd, dtype=np.float32)
        _check(lib().b2nn_set_normalisation(self._h, _host_ptr(mean), _host_ptr(sd), mean.size), "b2nn_set_normalisation")

    def get_normalisation(self):
        n = self.layers[0]
        mean, sd = np.zeros(n, dtype=np.float32), np.zeros(n, dtype=np.float32)
        _check(lib().b2nn_get_normalisation(self._h, _host_ptr(mean), _host_ptr(sd), n), "b2nn_get_normalisation")
        return mean, sd

    def save_checkpoint(self, path: str):
        _check(lib().b2nn_save_checkpoint(self._h, os.fsencode(path)), "b2nn_save_checkpoint")

    def load_checkpoint(self, path: str, layers=None):
        _check(lib().b2nn_load_checkpoint(self._h, os.fsencode(path)), "b2nn_load_checkpoint")
        if layers is not None:
            self.layers = list(layers)

    def evaluate_raw(self, desc: TrainDesc, x_raw_cm, y, rows: int, y_is_labels=False, normalise=True):
        x_raw_cm = np.ascontiguousarray(x_raw_cm, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        cost, errs = C.c_double(), C.c_double()
        _check(lib().b2nn_evaluate_raw(self._h, C.byref(desc), _host_ptr(x_raw_cm), _host_ptr(y), rows, int(y_is_labels),
                                       int(normalise), C.byref(cost), C.byref(errs)), "b2nn_evaluate_raw")
        return cost.value, errs.value

    def predict_raw(self, desc: TrainDesc, x_raw_cm, rows: int, normalise=True):
        x_raw_cm = np.ascontiguousarray(x_raw_cm, dtype=np.float32)
        K = self.layers[-1]
        out = np.empty(rows if desc.classify else rows * K, dtype=np.float32)
        probs = np.empty(rows * K, dtype=np.float32)
        _check(lib().b2nn_predict_raw(self._h, C.byref(desc), _host_ptr(x_raw_cm), rows, int(normalise), _host_ptr(out),
                                      _host_ptr(probs)), "b2nn_predict_raw")
        return out, probs

    def predict_ptr(self, desc: TrainDesc, x_ptr, rows: int, out_ptr, probs_ptr=None):
        """b2nn_predict on raw host pointers (e.g. pinned torch tensors): no numpy copies on either side."""
        _check(lib().b2nn_predict(self._h, C.byref(desc), x_ptr, rows, out_ptr, probs_ptr), "b2nn_predict")

    def train(self, desc: TrainDesc) -> TrainOut:
        its, js = [], []

        @LOG_CB
        def cb(_user, it, J):
            its.append(it)
            js.append(J)

        r = TrainResult()
        _check(lib().b2nn_train(self._h, C.byref(desc), C.cast(cb, C.c_void_p), None, C.byref(r)), "b2nn_train")
        return TrainOut(np.array(its, dtype=np.int32), np.array(js, dtype=np.float64), r.final_cost, r.error_count,
                        r.call_seconds, r.loop_seconds, r.steps, r.kernel_launches, r.gemm_tc_launches)

    def train_step_host(self, desc: TrainDesc, x_cm_ptr, y_cm_ptr, rows: int, global_rows: int = 0) -> float:
        J = C.c_float()
        _check(lib().b2nn_train_step_host(self._h, C.byref(desc), x_cm_ptr, y_cm_ptr, rows, global_rows, C.byref(J)),
               "b2nn_train_step_host")
        return float(J.value)

    def train_host(self, desc: TrainDesc, x_ptr, y_ptr, m: int):
        """Streaming trainer over host buffers (ctypes pointers).  Returns (per-step costs, TrainResult)."""
        steps = desc.iters * desc.batch_num
        J = np.empty(steps, dtype=np.float32)
        r = TrainResult()
        _check(lib().b2nn_train_host(self._h, C.byref(desc), x_ptr, y_ptr, m, self.layers[0] + 1, self.layers[-1],
                                     J.ctypes.data_as(C.POINTER(C.c_float)), C.byref(r)), "b2nn_train_host")
        return J, r

    def reset_velocity(self):
        _check(lib().b2nn_reset_velocity(self._h), "b2nn_reset_velocity")

    def evaluate(self, desc: TrainDesc, x_cm, y, rows: int, y_is_labels=False):
        x_cm = np.ascontiguousarray(x_cm, dtype=np.float32)
        y = np.ascontiguousarray(y, dtype=np.float32)
        cost, errs = C.c_double(), C.c_double()
        _check(lib().b2nn_evaluate(self._h, C.byref(desc), _host_ptr(x_cm), _host_ptr(y), rows, int(y_is_labels),
                                   C.byref(cost), C.byref(errs)), "b2nn_evaluate")
        return cost.value, errs.value

    def predict(self, desc: TrainDesc, x_cm, rows: int):
        x_cm = np.ascontiguousarray(x_cm, dtype=np.float32)
        K = self.layers[-1]
        out = np.empty(rows if desc.classify else rows * K, dtype=np.float32)
        probs = np.empty(rows * K, dtype=np.float32)
        _check(lib().b2nn_predict(self._h, C.byref(desc), _host_ptr(x_cm), rows, _host_ptr(out), _host_ptr(probs)),
               "b2nn_predict")
        return out, probs

    def predict_device(self, desc: TrainDesc, x_dev_ptr: int, rows: int, ld: int, out_dev_ptr: int):
        _check(lib().b2nn_predict_device(self._h, C.byref(desc), C.c_void_p(x_dev_ptr), rows, ld,
                                         C.c_void_p(out_dev_ptr)), "b2nn_predict_device")

    def gemm(self, backend, A, lda, a_major, B, ldb, b_major, D, ldd, M, N, K, epi=EPI_STORE, aux=0, ldaux=0):
        """Raw contraction on DEVICE pointers (ints)."""
        _check(lib().b2nn_gemm(self._h, backend, C.c_void_p(A), lda, a_major, C.c_void_p(B), ldb, b_major,
                               C.c_void_p(D), ldd, M, N, K, epi, C.c_void_p(aux) if aux else None, ldaux), "b2nn_gemm")

    def momentum_update(self, theta, vel, grad, count, mu, alpha):
        _check(lib().b2nn_momentum_update(self._h, C.c_void_p(theta), C.c_void_p(vel), C.c_void_p(grad), count, mu,
                                          alpha), "b2nn_momentum_update")

    def stream(self) -> int:
        return int(lib().b2nn_stream(self._h) or 0)

    def profile_enable(self, on: bool):
        _check(lib().b2nn_profile_enable(self._h, int(on)), "b2nn_profile_enable")

    def profile_read(self) -> dict:
        n = len(PROFILE_KINDS)
        ms, fl, by = (C.c_double * n)(), (C.c_double * n)(), (C.c_double * n)()
        la, tc = (C.c_int64 * n)(), (C.c_int64 * n)()
        _check(lib().b2nn_profile_read(self._h, ms, fl, by, la, tc), "b2nn_profile_read")
        return {k: dict(ms=ms[i], flops=fl[i], bytes=by[i], launches=int(la[i]), tc_launches=int(tc[i]))
                for i, k in enumerate(PROFILE_KINDS)}

    # ---- data parallel ----
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(128)
        _check(lib().b2nn_comm_unique_id(buf), "b2nn_comm_unique_id")
        return buf.raw

    def comm_init(self, rank: int, nranks: int, uid: bytes):
        buf = C.create_string_buffer(uid, 128)
        _check(lib().b2nn_comm_init(self._h, rank, nranks, buf), "b2nn_comm_init")

    def comm_p2p_export(self) -> bytes:
        buf = C.create_string_buffer(192)
        _check(lib().b2nn_comm_p2p_export(self._h, buf), "b2nn_comm_p2p_export")
        return buf.raw

    def comm_p2p_open(self, blobs):
        """blobs: the world x 192 bytes gathered in rank order; None switches the peer-memory path off again."""
        buf = C.create_string_buffer(blobs, len(blobs)) if blobs is not None else None
        _check(lib().b2nn_comm_p2p_open(self._h, buf), "b2nn_comm_p2p_open")

    def comm_barrier(self):
        _check(lib().b2nn_comm_barrier(self._h), "b2nn_comm_barrier")
