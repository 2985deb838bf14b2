"""The drop-in surface: the reference's own main.cpp must compile UNMODIFIED against include/ (CPU check, only where
/root/reference is mounted), the examples restating its commented-out experiments must build and — on the GPU — produce
the oracle's numbers through class cublasNN."""
import os
import re
import shutil
import subprocess

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "cublas-ml_b200", "lib")
CXX = "/usr/bin/g++"


def _build(src, out, std="c++11"):
    cmd = [CXX, "-O1", f"-std={std}", "-w", "-I", os.path.join(ROOT, "include"), src, "-o", out, "-L", LIB,
           "-lcublasNN_b200", "-lb2nn", f"-Wl,-rpath,{LIB}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-3000:]


def test_reference_main_compiles_unmodified(tmp_path):
    ref = "/root/reference/main.cpp"
    if not os.path.exists(ref):
        pytest.skip("/root/reference is not mounted here")
    if not os.path.exists(os.path.join(LIB, "libcublasNN_b200.so")):
        pytest.skip("facade not built")
    # a copy in an empty directory, so `#include "cublasNN.h"` can only resolve to include/cublasNN.h
    dst = tmp_path / "main.cpp"
    shutil.copyfile(ref, dst)
    _build(str(dst), str(tmp_path / "main_ref"))
    assert os.path.getsize(tmp_path / "main_ref") > 0


@pytest.mark.parametrize("name", ["bikeshare_example", "mnist_example"])
def test_examples_build(tmp_path, name):
    if not os.path.exists(os.path.join(LIB, "libcublasNN_b200.so")):
        pytest.skip("facade not built")
    _build(os.path.join(ROOT, "examples", name + ".cpp"), str(tmp_path / name))


def _write_csv(path, a, fmt):
    np.savetxt(path, a, delimiter=",", fmt=fmt)


def test_facade_program_fails_loudly_without_a_gpu(tmp_path, b2nn):
    """No sm_100 device: a program written against class cublasNN must stop with a message and a non-zero status — never
    run to the end on a CPU path or print results it did not compute."""
    if b2nn.lib().b2nn_device_ok():
        pytest.skip("a usable GPU is present")
    if not os.path.exists(os.path.join(LIB, "libcublasNN_b200.so")):
        pytest.skip("facade not built")
    rng = np.random.default_rng(0)
    _write_csv(tmp_path / "trainX.csv", rng.integers(0, 255, size=(40, 784)), "%d")
    _write_csv(tmp_path / "trainY.csv", rng.integers(0, 10, size=(40, 1)), "%d")
    exe = tmp_path / "mnist_example"
    _build(os.path.join(ROOT, "examples", "mnist_example.cpp"), str(exe))
    r = subprocess.run([str(exe), str(tmp_path / "trainX.csv"), str(tmp_path / "trainY.csv"), "2", "5", "42"],
                       capture_output=True, text=True, cwd=tmp_path, timeout=120)
    assert r.returncode != 0, r.stdout[-1000:]
    assert "b2nn: no" in r.stderr, r.stderr[-1000:]
    assert "Final Cost" not in r.stdout and "errors" not in r.stdout, r.stdout[-1000:]


@pytest.mark.gpu
def test_bikeshare_example_matches_oracle(tmp_path, oracle):
    """Config C2 end to end through the class API (CSV -> normalise -> train -> validate -> predict), fp32 mode,
    against the CPU oracle fed with the same normalised data and the same cuRAND initial weights."""
    b = cases.bikeshare()
    d = tmp_path / "bikeshare"
    d.mkdir()
    def with_col1(x):       # the example erases column 1 again (main.cpp:66-71): re-insert a dummy
        return np.insert(x, 1, 0.0, axis=1)
    _write_csv(d / "trainP7_1.csv", with_col1(b["x"]), "%.9g")
    _write_csv(d / "trainYP2_1.csv", b["y"], "%.9g")
    _write_csv(d / "validateP7_1.csv", with_col1(b["xv"]), "%.9g")
    _write_csv(d / "validateY1.csv", b["yv"], "%.9g")
    _write_csv(d / "testPF1_1.csv", with_col1(b["xp"]), "%.9g")
    exe = tmp_path / "bikeshare_example"
    _build(os.path.join(ROOT, "examples", "bikeshare_example.cpp"), str(exe))
    iters, seed = 150, 42       # stops before the trajectory starts to oscillate (around iteration 180 at rate 0.003)
    env = dict(os.environ, B2NN_EXAMPLE_FP32="1")
    r = subprocess.run([str(exe), str(d), str(iters), str(seed)], capture_output=True, text=True, cwd=tmp_path, env=env,
                       timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    out = r.stdout
    its = [(int(a), float(c)) for a, c in re.findall(r"Iteration: (\d+)\tCost: ([-+0-9.eE]+)", out)]
    final = float(re.search(r"Final Cost: ([-+0-9.eE]+)", out).group(1))
    val = float(re.search(r"Final Validation Cost: ([-+0-9.eE]+)", out).group(1))
    assert "rows predicted: 3223" in out

    # oracle on the same problem: float32 z-score with ddof = 1 like normaliseData (cublasNN.cpp:191-268)
    x = b["x"].astype(np.float32)
    mu = x.sum(axis=0, dtype=np.float32) / np.float32(len(x))
    sd = np.sqrt(((x - mu) ** 2).sum(axis=0, dtype=np.float32) / np.float32(len(x) - 1)).astype(np.float32)
    xn = ((x - mu) / sd).astype(np.float32)
    layers = [10, 40, 160, 1]
    th0 = oracle.init_weights(layers, 1, seed)
    ref = oracle.train(layers, oracle.with_bias_colmajor(xn), oracle.colmajor(b["y"]), len(x), batch_num=1, iters=iters,
                       display=iters // 10, momentum=0.9, rate=0.003, hidden_act=oracle.SIGMOID, classify=False,
                       theta0=th0, precision="f64")
    assert [i for i, _ in its] == list(ref.log_iter)
    got = np.array([c for _, c in its])
    np.testing.assert_allclose(got, ref.log_J, rtol=1e-4)               # the class prints 6 significant digits
    assert abs(final - ref.final_cost) <= 1e-3 * ref.final_cost
    assert val > 0


@pytest.mark.gpu
def test_mnist_example_runs_and_learns(tmp_path):
    lab = cases.mnist_labels()[:1500]
    x = cases.synth_digits(lab, 784, seed=1234)
    _write_csv(tmp_path / "trainX.csv", x, "%d")
    _write_csv(tmp_path / "trainY.csv", lab.reshape(-1, 1), "%d")
    exe = tmp_path / "mnist_example"
    _build(os.path.join(ROOT, "examples", "mnist_example.cpp"), str(exe))
    r = subprocess.run([str(exe), str(tmp_path / "trainX.csv"), str(tmp_path / "trainY.csv"), "10", "5", "42"],
                       capture_output=True, text=True, cwd=tmp_path, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    costs = [float(c) for c in re.findall(r"Iteration: \d+\tCost: ([-+0-9.eE]+)", r.stdout)]
    assert len(costs) == 10 and costs[-1] < 0.1 * costs[0]
    errs = float(re.search(r"Total errors \t([-+0-9.eE]+)", r.stdout).group(1))
    wrong = int(re.search(r"predict errors on the training set: (\d+)", r.stdout).group(1))
    assert errs <= 15 and wrong == int(errs)
