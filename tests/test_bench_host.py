"""CPU checks of bench.py's host logic: the FLOP model of SURVEY.md 8(d), the traffic gate (roofline.traffic is quoted
only for the kernel source that was profiled) and the argument contract the driver relies on."""
import hashlib
import importlib.util
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_under_test", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_step_flops_is_the_survey_model(bench):
    # 2 * sum_j 2 B (l_j + 1) l_{j+1}  (forward + weight gradient)  +  sum_{j>=1} 2 B l_j l_{j+1}  (delta backprop)
    layers, B = [784, 4096, 4096, 10], 8192
    fwd = sum(2 * B * (a + 1) * b for a, b in zip(layers[:-1], layers[1:]))
    bwd = sum(2 * B * a * b for a, b in zip(layers[1:-1], layers[2:]))
    assert bench.step_flops(layers, B) == 2 * fwd + bwd == 932142448640


def test_traffic_is_quoted_only_for_the_profiled_kernel_source(bench, tmp_path):
    b2nn_dir = os.path.join(ROOT, "cublas-ml_b200")
    rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    sha = hashlib.sha256(open(os.path.join(b2nn_dir, "csrc", "gemm_tc.cu"), "rb").read()).hexdigest()[:16]
    value, src = bench.traffic_record(b2nn_dir)
    if rec["gemm_tc_cu_sha16"] == sha:
        # the committed capture is of the shipped kernel: the bench may (and does) quote it
        assert value == rec["gemm_tc_kernel"]["dram_bytes_per_launch_mean"] and src["profiled"].startswith("profiles/r2_traffic.json@")
    else:
        assert value is None and src["profiled"] is None
    # a different kernel source: never quoted
    fake = tmp_path / "pkg" / "csrc"
    fake.mkdir(parents=True)
    (fake / "gemm_tc.cu").write_text("// not the profiled kernel\n")
    value, src = bench.traffic_record(str(tmp_path / "pkg"))
    assert value is None and src["profiled"] is None


def test_committed_capture_is_of_the_shipped_kernel():
    """profiles/r2_traffic.json must describe the gemm_tc.cu in the tree (re-capture with tools/ncu_summarise.py after a
    kernel change)."""
    rec = json.load(open(os.path.join(ROOT, "profiles", "r2_traffic.json")))
    sha = hashlib.sha256(open(os.path.join(ROOT, "cublas-ml_b200", "csrc", "gemm_tc.cu"), "rb").read()).hexdigest()[:16]
    assert rec["gemm_tc_cu_sha16"] == sha


def test_bench_fails_loudly_without_a_gpu(b2nn):
    if b2nn.lib().b2nn_device_ok():
        pytest.skip("a usable GPU is present")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--steps", "2", "--warmup", "1", "--no-secondary"],
                       capture_output=True, text=True, cwd=ROOT, timeout=300)
    assert r.returncode != 0
    assert '"value"' not in r.stdout, r.stdout[-500:]
