"""ctypes bindings for the TEST-INFRASTRUCTURE libraries under oracle/.

* ``liboracle.so``              — CPU restatement of the reference hot path (oracle/mlp_oracle.cpp).
* ``_ref/libcublasml_ref.so``   — the unmodified reference built from /root/reference (oracle/ref_harness.cpp);
                                  needs a GPU, only usable on the B200 box.

Only tests/, bench.py's cpu_baseline / ``--impl reference`` legs and __graft_entry__.smoke() may import this
module.  Product code never does.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))

SIGMOID, TANH = 0, 1
OUT_SIGMOID, OUT_SOFTMAX = 0, 1
COST_NEGLNMAX, COST_CROSSENTROPY = 0, 1


class _Problem(C.Structure):
    _fields_ = [("layers", C.POINTER(C.c_int)), ("n_layers", C.c_int), ("x", C.POINTER(C.c_float)),
                ("y", C.POINTER(C.c_float)), ("m", C.c_int64), ("batch_num", C.c_int), ("iters", C.c_int),
                ("display", C.c_int), ("momentum", C.c_float), ("rate", C.c_float), ("hidden_act", C.c_int),
                ("classify", C.c_int), ("out_act", C.c_int), ("cost", C.c_int)]


class _Result(C.Structure):
    _fields_ = [("final_cost", C.c_double), ("error_count", C.c_double), ("n_logged", C.c_int),
                ("loop_seconds", C.c_double)]


class _RefProblem(C.Structure):
    _fields_ = [("layers", C.POINTER(C.c_int)), ("n_layers", C.c_int), ("x", C.POINTER(C.c_float)),
                ("y", C.POINTER(C.c_float)), ("m", C.c_int64), ("n", C.c_int), ("ycols", C.c_int),
                ("batch_num", C.c_int), ("iters", C.c_int), ("display", C.c_int), ("momentum", C.c_float),
                ("rate", C.c_float), ("hidden_act", C.c_int), ("classify", C.c_int), ("out_act", C.c_int),
                ("cost", C.c_int), ("init_kind", C.c_int)]


class _RefResult(C.Structure):
    _fields_ = [("final_cost", C.c_double), ("error_count", C.c_double), ("n_logged", C.c_int),
                ("call_seconds", C.c_double), ("wall_seconds", C.c_double)]


@dataclass
class TrainOut:
    theta: np.ndarray
    log_iter: np.ndarray
    log_J: np.ndarray
    final_cost: float
    error_count: float
    seconds: float
    theta0: np.ndarray | None = None
    extra: dict = field(default_factory=dict)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: run `make -C oracle` (or __graft_entry__.build())")
        _lib = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        for name in ("oracle_train_f32", "oracle_train_f64"):
            f = getattr(_lib, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(_Problem), fp, fp, C.POINTER(C.c_int), C.POINTER(C.c_double), C.c_int,
                          C.POINTER(_Result)]
        for name in ("oracle_predict_f32", "oracle_predict_f64"):
            f = getattr(_lib, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(C.c_int), C.c_int, fp, fp, C.c_int64, C.c_int, C.c_int, C.c_int, fp, fp]
        _lib.oracle_theta_size.restype = C.c_int64
        _lib.oracle_theta_size.argtypes = [C.POINTER(C.c_int), C.c_int]
        _lib.oracle_init_weights.restype = C.c_int
        _lib.oracle_init_weights.argtypes = [fp, C.POINTER(C.c_int), C.c_int, C.c_int, C.c_ulonglong]
        _lib.oracle_scale_uniform.restype = None
        _lib.oracle_scale_uniform.argtypes = [fp, C.POINTER(C.c_int), C.c_int, C.c_int]
        _lib.oracle_num_threads.restype = C.c_int
    return _lib


def ref_available() -> bool:
    return os.path.exists(os.path.join(_HERE, "_ref", "libcublasml_ref.so"))


def ref():
    """The unmodified reference (cuBLAS build).  Loading works anywhere; calling needs a GPU."""
    global _ref
    if _ref is None:
        path = os.path.join(_HERE, "_ref", "libcublasml_ref.so")
        if not os.path.exists(path):
            raise RuntimeError(f"{path} missing: `make -C oracle ref` needs /root/reference (dev container only)")
        _ref = C.CDLL(path)
        fp = C.POINTER(C.c_float)
        _ref.ref_train.restype = C.c_int
        _ref.ref_train.argtypes = [C.POINTER(_RefProblem), fp, C.c_ulonglong, fp, fp, C.POINTER(C.c_int),
                                   C.POINTER(C.c_double), C.c_int, C.POINTER(_RefResult)]
        _ref.ref_predict.restype = C.c_int
        _ref.ref_predict.argtypes = [C.POINTER(_RefProblem), fp, fp]
        _ref.ref_device_count.restype = C.c_int
    return _ref


def _fptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_float)) if a is not None else None


def _iptr(a):
    return a.ctypes.data_as(C.POINTER(C.c_int))


def theta_size(layers) -> int:
    l = np.ascontiguousarray(layers, dtype=np.int32)
    return int(lib().oracle_theta_size(_iptr(l), len(l)))


def theta_offsets(layers):
    """Per-layer (offset, rows = l_j+1, cols = l_{j+1}) in the reference's arena (cublasNN.cpp:293-301)."""
    out, pos = [], 0
    for j in range(len(layers) - 1):
        r, c = layers[j] + 1, layers[j + 1]
        out.append((pos, r, c))
        pos += r * c
    return out


def init_weights(layers, init_kind: int, seed: int) -> np.ndarray:
    """setLayers' initial state: cuRAND XORWOW uniform over the arena, then the rescale kernel."""
    l = np.ascontiguousarray(layers, dtype=np.int32)
    th = np.empty(theta_size(layers), dtype=np.float32)
    rc = lib().oracle_init_weights(_fptr(th), _iptr(l), len(l), init_kind, seed)
    if rc != 0:
        raise RuntimeError(f"oracle_init_weights failed rc={rc}")
    return th


def with_bias_colmajor(x_rows: np.ndarray) -> np.ndarray:
    """vector2dToMat + addBias (cublasNN.cpp:116-133, 330-346): rows of features -> column-major m x (n+1)
    with a leading ones column, returned as a 1-D float32 buffer."""
    m, n = x_rows.shape
    out = np.empty((n + 1, m), dtype=np.float32)   # row k of this array == column k of the col-major matrix
    out[0, :] = 1.0
    out[1:, :] = x_rows.T
    return np.ascontiguousarray(out).reshape(-1)


def colmajor(a_rows: np.ndarray) -> np.ndarray:
    return np.ascontiguousarray(a_rows.T.astype(np.float32)).reshape(-1)


def one_hot_colmajor(labels: np.ndarray, K: int) -> np.ndarray:
    """classToBin (cublasNN.cpp:135-146)."""
    m = labels.shape[0]
    y = np.zeros((K, m), dtype=np.float32)
    y[labels.astype(np.int64), np.arange(m)] = 1.0
    return y.reshape(-1)


def train(layers, x_cm, y_cm, m, *, batch_num=1, iters=1, display=1, momentum=0.9, rate=0.01, hidden_act=TANH,
          classify=True, out_act=OUT_SOFTMAX, cost=COST_CROSSENTROPY, theta0=None, precision="f32") -> TrainOut:
    """Run the CPU oracle.  x_cm: column-major m x (l0+1) incl. ones column; y_cm: column-major m x l_last."""
    l = np.ascontiguousarray(layers, dtype=np.int32)
    x_cm = np.ascontiguousarray(x_cm, dtype=np.float32)
    y_cm = np.ascontiguousarray(y_cm, dtype=np.float32)
    assert x_cm.size == m * (layers[0] + 1), (x_cm.size, m, layers[0])
    assert y_cm.size == m * layers[-1]
    theta0 = np.ascontiguousarray(theta0, dtype=np.float32)
    assert theta0.size == theta_size(layers)
    p = _Problem(_iptr(l), len(l), _fptr(x_cm), _fptr(y_cm), m, batch_num, iters, display, momentum, rate,
                 hidden_act, int(classify), out_act, cost)
    cap = iters // max(display, 1) + 2
    li = np.zeros(cap, dtype=np.int32)
    lj = np.zeros(cap, dtype=np.float64)
    th = np.empty_like(theta0)
    r = _Result()
    fn = lib().oracle_train_f32 if precision == "f32" else lib().oracle_train_f64
    rc = fn(C.byref(p), _fptr(theta0), _fptr(th), _iptr(li), lj.ctypes.data_as(C.POINTER(C.c_double)), cap,
            C.byref(r))
    if rc != 0:
        raise RuntimeError(f"oracle train failed rc={rc}")
    n = r.n_logged
    return TrainOut(th, li[:n].copy(), lj[:n].copy(), r.final_cost, r.error_count, r.loop_seconds)


def predict(layers, theta, x_cm, rows, *, hidden_act=TANH, classify=True, out_act=OUT_SOFTMAX, precision="f32"):
    """Forward only.  Returns (out, probs): out = class ids (rows,) or col-major rows x K targets."""
    l = np.ascontiguousarray(layers, dtype=np.int32)
    theta = np.ascontiguousarray(theta, dtype=np.float32)
    x_cm = np.ascontiguousarray(x_cm, dtype=np.float32)
    K = layers[-1]
    out = np.empty(rows if classify else rows * K, dtype=np.float32)
    probs = np.empty(rows * K, dtype=np.float32)
    fn = lib().oracle_predict_f32 if precision == "f32" else lib().oracle_predict_f64
    rc = fn(_iptr(l), len(l), _fptr(theta), _fptr(x_cm), rows, hidden_act, int(classify), out_act, _fptr(out),
            _fptr(probs))
    if rc != 0:
        raise RuntimeError(f"oracle predict failed rc={rc}")
    return out, probs


def ref_train(layers, x_rows, y_rows, *, batch_num=1, iters=1, display=1, momentum=0.9, rate=0.01,
              hidden_act=TANH, classify=True, out_act=OUT_SOFTMAX, cost=COST_CROSSENTROPY, init_kind=2, seed=42,
              theta0=None) -> TrainOut:
    """Run the UNMODIFIED reference on the GPU.  x_rows: (m, n) float32 rows; y_rows: (m, 1) labels when classify,
    (m, K) targets otherwise — i.e. what readCSV would hand to setData."""
    l = np.ascontiguousarray(layers, dtype=np.int32)
    x_rows = np.ascontiguousarray(x_rows, dtype=np.float32)
    y_rows = np.ascontiguousarray(y_rows, dtype=np.float32).reshape(x_rows.shape[0], -1)
    m, n = x_rows.shape
    p = _RefProblem(_iptr(l), len(l), _fptr(x_rows), _fptr(y_rows), m, n, y_rows.shape[1], batch_num, iters,
                    display, momentum, rate, hidden_act, int(classify), out_act, cost, init_kind)
    cap = iters // max(display, 1) + 2
    li = np.zeros(cap, dtype=np.int32)
    lj = np.zeros(cap, dtype=np.float64)
    nt = theta_size(layers)
    th0 = np.empty(nt, dtype=np.float32)
    th = np.empty(nt, dtype=np.float32)
    t0 = None if theta0 is None else np.ascontiguousarray(theta0, dtype=np.float32)
    r = _RefResult()
    rc = ref().ref_train(C.byref(p), _fptr(t0), seed, _fptr(th0), _fptr(th), _iptr(li),
                         lj.ctypes.data_as(C.POINTER(C.c_double)), cap, C.byref(r))
    if rc != 0:
        raise RuntimeError(f"ref_train failed rc={rc}")
    k = r.n_logged
    return TrainOut(th, li[:k].copy(), lj[:k].copy(), r.final_cost, r.error_count, r.call_seconds, theta0=th0,
                    extra={"wall_seconds": r.wall_seconds})


def ref_predict(layers, theta, x_rows, *, hidden_act=TANH, classify=True, out_act=OUT_SOFTMAX):
    l = np.ascontiguousarray(layers, dtype=np.int32)
    x_rows = np.ascontiguousarray(x_rows, dtype=np.float32)
    theta = np.ascontiguousarray(theta, dtype=np.float32)
    m, n = x_rows.shape
    y = np.zeros(1, dtype=np.float32)
    p = _RefProblem(_iptr(l), len(l), _fptr(x_rows), _fptr(y), m, n, 1, 1, 1, 1, 0.0, 0.0, hidden_act,
                    int(classify), out_act, 0, 1)
    cols = 1 if classify else layers[-1]
    out = np.empty(m * cols, dtype=np.float32)
    rc = ref().ref_predict(C.byref(p), _fptr(theta), _fptr(out))
    if rc != 0:
        raise RuntimeError(f"ref_predict failed rc={rc}")
    return out.reshape(m, cols)
