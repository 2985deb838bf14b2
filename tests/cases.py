"""Seeded parity cases shared by the golden-vector generator (tests/golden/make_golden.py) and the tests.

Every case is a small instance of the reference's two entry points, trainClassifyMomentum / trainFuncApproxMomentum
(cublasNN.cpp:526-640), sized so the CPU oracle finishes in seconds.  Inputs come from numpy's PCG64 (stable across
numpy versions) or from the committed data fixtures; nothing here reads /root/reference.
"""
from __future__ import annotations

import functools
import os
from dataclasses import dataclass, field

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
GOLDEN = os.path.join(HERE, "golden")

SIGMOID, TANH = 0, 1
OUT_SIGMOID, OUT_SOFTMAX = 0, 1
COST_NEGLNMAX, COST_CROSSENTROPY = 0, 1


@dataclass
class Case:
    name: str
    layers: list
    x: np.ndarray                 # (m, n) float32 rows, already normalised, NO bias column
    y: np.ndarray                 # (m, 1) class ids when classify, (m, K) targets otherwise
    classify: bool
    hidden_act: int
    out_act: int = OUT_SOFTMAX
    cost: int = COST_CROSSENTROPY
    batch_num: int = 1
    iters: int = 10
    display: int = 1
    momentum: float = 0.9
    rate: float = 0.05
    init_kind: int = 2
    seed: int = 42
    compare_log: bool = True      # False when m % batch_num != 0: the reference's printed J then includes stale rows (Q2)
    notes: str = ""
    extra: dict = field(default_factory=dict)

    @property
    def m(self):
        return self.x.shape[0]


def zscore(x: np.ndarray) -> np.ndarray:
    """normaliseData (cublasNN.cpp:191-212): (x - mean) / stddev(ddof=1), zero where the column is constant."""
    x = x.astype(np.float64)
    mu = x.mean(axis=0)
    sd = x.std(axis=0, ddof=1)
    out = np.where(sd != 0, (x - mu) / np.where(sd != 0, sd, 1.0), 0.0)
    return out.astype(np.float32)


def synth_digits(labels: np.ndarray, n_features: int, seed: int) -> np.ndarray:
    """C1 stand-in pixels (the MNIST pixel CSVs are absent from the reference checkout): one sparse
    prototype per class plus noise, clipped to 0..255 like 8-bit pixels."""
    rng = np.random.Generator(np.random.PCG64(seed))
    K = int(labels.max()) + 1
    proto = rng.integers(0, 256, size=(K, n_features)).astype(np.float32)
    proto *= (rng.random((K, n_features)) < 0.19)
    noise = rng.standard_normal((labels.shape[0], n_features)).astype(np.float32) * 48.0
    x = np.clip(np.rint(proto[labels] + noise), 0, 255)
    return x.astype(np.float32)


def mnist_labels() -> np.ndarray:
    return np.load(os.path.join(GOLDEN, "mnist_labels.npy")).astype(np.int64)


def bikeshare():
    return np.load(os.path.join(GOLDEN, "bikeshare.npz"))


def _clf_blobs(m, n, K, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    centers = rng.standard_normal((K, n)).astype(np.float32) * 1.5
    lab = rng.integers(0, K, size=m)
    x = centers[lab] + rng.standard_normal((m, n)).astype(np.float32)
    return zscore(x), lab.astype(np.float32).reshape(m, 1)


@functools.lru_cache(maxsize=1)
def small_cases() -> list[Case]:
    cases = []
    x, y = _clf_blobs(300, 20, 5, seed=1)
    cases.append(Case("clf_tanh_softmax", [20, 16, 12, 5], x, y, True, TANH, OUT_SOFTMAX, COST_CROSSENTROPY,
                      batch_num=4, iters=30, display=5, rate=0.05, init_kind=2, seed=42))
    x, y = _clf_blobs(257, 13, 4, seed=2)
    cases.append(Case("clf_sigmoid_sigout", [13, 9, 4], x, y, True, SIGMOID, OUT_SIGMOID, COST_NEGLNMAX,
                      batch_num=1, iters=40, display=10, rate=0.5, init_kind=1, seed=7))
    b = bikeshare()
    cases.append(Case("reg_sigmoid_bikeshare512", [10, 40, 160, 1], zscore(b["x"][:512]), b["y"][:512].astype(np.float32),
                      False, SIGMOID, batch_num=1, iters=200, display=50, rate=0.003, init_kind=1, seed=42,
                      notes="first 512 rows of config C2 (main.cpp:54-96)"))
    rng = np.random.Generator(np.random.PCG64(3))
    x = zscore(rng.standard_normal((205, 6)).astype(np.float32))
    w = rng.standard_normal((6, 3)).astype(np.float32)
    y = np.tanh(x @ w) + 0.05 * rng.standard_normal((205, 3)).astype(np.float32)
    cases.append(Case("reg_tanh_ragged", [6, 10, 3], x, y.astype(np.float32), False, TANH, batch_num=7, iters=25,
                      display=5, rate=0.02, init_kind=2, seed=11, compare_log=False,
                      notes="m % batch_num != 0: last batch absorbs the remainder (cublasNN.cpp:429)"))
    lab = mnist_labels()[:2000]
    cases.append(Case("mnist_mini", [784, 50, 10], zscore(synth_digits(lab, 784, seed=1234)),
                      lab.astype(np.float32).reshape(-1, 1), True, TANH, OUT_SOFTMAX, COST_CROSSENTROPY, batch_num=10,
                      iters=5, display=1, rate=0.075, init_kind=2, seed=42,
                      notes="config C1 (main.cpp:126-157) on the first 2000 labels"))
    x, y = _clf_blobs(2048, 256, 10, seed=5)
    cases.append(Case("wide_mini", [256, 512, 384, 10], x, y, True, TANH, OUT_SOFTMAX, COST_CROSSENTROPY, batch_num=2,
                      iters=3, display=1, rate=0.02, init_kind=2, seed=42,
                      notes="scaled-down config C3: every contraction is large enough for the tcgen05 path"))
    return cases


def extra_cases() -> list:
    """Cases without golden vectors of the reference (checked against the oracle / a single-GPU run only)."""
    x, y = _clf_blobs(4096, 128, 10, seed=9)
    return [Case("wide1024_skinny", [128, 1024, 10], x, y, True, TANH, OUT_SOFTMAX, COST_CROSSENTROPY, batch_num=2, iters=3,
                 display=1, rate=0.02, init_kind=2, seed=7,
                 notes="1024 hidden units in front of 10 outputs: fused delta-backprop + gradient kernel (skinny.cu); data "
                       "parallel: layer 0 splits into 256-row tiles per rank, the output layer goes through slot buffers")]


@functools.lru_cache(maxsize=1)
def c4_mini() -> Case:
    """Scaled-down config C4 (784 - 8 x 8192 - 10): nine weight layers, widths chosen so that with 2, 4 and 8 ranks some
    layers split into whole 256-column tiles per rank (fused peer-memory exchange inside the gradient GEMM) and others
    do not (slot buffers); 4096-row batches keep every rank's contractions on the tcgen05 path up to 8 ranks."""
    layers = [64, 512, 256, 1024, 512, 2048, 256, 512, 1024, 10]
    x, y = _clf_blobs(8192, 64, 10, seed=13)
    return Case("c4_mini", layers, x, y, True, TANH, OUT_SOFTMAX, COST_CROSSENTROPY, batch_num=2, iters=2, display=1,
                rate=0.02, init_kind=2, seed=5, notes="config C4 in miniature: 9 weight layers, data-parallel schedule")


def case_by_name(name: str) -> Case:
    if name == "c4_mini":
        return c4_mini()
    for c in small_cases() + extra_cases():
        if c.name == name:
            return c
    raise KeyError(name)


def golden_path(name: str) -> str:
    return os.path.join(GOLDEN, f"ref_{name}.npz")


def oracle_inputs(c: Case):
    """(x_cm, y_cm) in the reference's host layout: column-major, ones column first / one-hot targets."""
    import sys
    root = os.path.dirname(HERE)
    if root not in sys.path:
        sys.path.insert(0, root)
    from oracle import oracle as O
    x_cm = O.with_bias_colmajor(c.x)
    if c.classify:
        y_cm = O.one_hot_colmajor(c.y[:, 0].astype(np.int64), c.layers[-1])
    else:
        y_cm = O.colmajor(c.y)
    return x_cm, y_cm
