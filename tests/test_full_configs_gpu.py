"""BASELINE.json configs C1 and C2 at FULL size: this engine against the UNMODIFIED reference (oracle/_ref, cuBLAS path)
run live on the same GPU with the same seed and data, and against the fp64 CPU oracle where it is fast enough.
C1: 784-50-10, m = 42000 real labels + synthetic pixels, 10 batches, tanh / softmax / cross-entropy, mu 0.9, rate 0.075
    (main.cpp:126-157); C2: bikeshare 10-40-160-1, m = 4337, full batch, sigmoid, mu 0.9, rate 0.003 (main.cpp:54-96)."""
import numpy as np
import pytest

import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu


def _engine_run(b2nn, layers, x_cm, y_cm, m, desc, init_kind, seed):
    net = b2nn.Net()
    net.set_layers(layers, init_kind, seed)
    th0 = net.get_weights()
    net.set_train_data(x_cm, y_cm, m)
    out = net.train(desc)
    th = net.get_weights()
    net.close()
    return th0, th, out


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3"])
def test_c1_full_size_against_reference_live(b2nn, oracle, prec):
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libcublasml_ref.so did not travel")
    layers, m, nb, iters = [784, 50, 10], 42000, 10, 30
    lab = cases.mnist_labels()
    x = cases.zscore(cases.synth_digits(lab, 784, seed=1234))
    y_rows = lab.astype(np.float32).reshape(-1, 1)
    r = oracle.ref_train(layers, x, y_rows, batch_num=nb, iters=iters, display=1, momentum=0.9, rate=0.075,
                         hidden_act=oracle.TANH, classify=True, out_act=oracle.OUT_SOFTMAX, cost=oracle.COST_CROSSENTROPY,
                         init_kind=2, seed=42)
    x_cm, y_cm = oracle.with_bias_colmajor(x), oracle.one_hot_colmajor(lab, 10)
    d = b2nn.make_desc(momentum=0.9, rate=0.075, iters=iters, display=1, batch_num=nb, classify=True, act=b2nn.ACT_TANH,
                       out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision={"fp32": 0, "tf32": 1, "tf32x3": 3}[prec])
    th0, th, out = _engine_run(b2nn, layers, x_cm, y_cm, m, d, b2nn.INIT_2, 42)
    assert np.array_equal(th0, r.theta0)
    assert list(out.log_iter) == list(r.log_iter)
    jtol = {"fp32": 2e-3, "tf32x3": 2e-3, "tf32": 5e-2}[prec]
    jerr = float(np.max(np.abs(out.log_J - r.log_J) / np.maximum(np.abs(r.log_J), 1e-6)))
    assert jerr <= jtol, (prec, jerr)
    assert abs(out.final_cost - r.final_cost) <= jtol * abs(r.final_cost) + 1e-6
    assert abs(out.error_count - r.error_count) <= 0.001 * m + 1          # +-0.1 % of m
    terr = float(np.abs(th - r.theta).max() / np.abs(r.theta).max())
    assert terr <= {"fp32": 2e-3, "tf32x3": 2e-3, "tf32": 1e-1}[prec], (prec, terr)


@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_c1_secondary_variant_full_batch_against_reference_live(b2nn, oracle, prec):
    """main.cpp:156 — the same net with sigmoid hidden units, sigmoid output and the negLnMax cost, batchNum 1 (one
    42 000-row batch per step)."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libcublasml_ref.so did not travel")
    layers, m, iters = [784, 50, 10], 42000, 12
    lab = cases.mnist_labels()
    x = cases.zscore(cases.synth_digits(lab, 784, seed=1234))
    kw = dict(batch_num=1, iters=iters, display=2, momentum=0.9, rate=0.075)
    r = oracle.ref_train(layers, x, lab.astype(np.float32).reshape(-1, 1), hidden_act=oracle.SIGMOID, classify=True,
                         out_act=oracle.OUT_SIGMOID, cost=oracle.COST_NEGLNMAX, init_kind=2, seed=42, **kw)
    x_cm, y_cm = oracle.with_bias_colmajor(x), oracle.one_hot_colmajor(lab, 10)
    d = b2nn.make_desc(classify=True, act=b2nn.ACT_SIGMOID, out=b2nn.OUT_SIGMOID, cost=b2nn.COST_NEGLNMAX,
                       precision={"fp32": 0, "tf32": 1}[prec], **kw)
    th0, th, out = _engine_run(b2nn, layers, x_cm, y_cm, m, d, b2nn.INIT_2, 42)
    assert np.array_equal(th0, r.theta0)
    assert list(out.log_iter) == list(r.log_iter)
    jtol = {"fp32": 2e-3, "tf32": 5e-2}[prec]
    np.testing.assert_allclose(out.log_J, r.log_J, rtol=jtol)
    assert abs(out.final_cost - r.final_cost) <= jtol * abs(r.final_cost) + 1e-6
    assert abs(out.error_count - r.error_count) <= 0.001 * m + 1
    terr = float(np.abs(th - r.theta).max() / np.abs(r.theta).max())
    assert terr <= {"fp32": 2e-3, "tf32": 1e-1}[prec], (prec, terr)


def test_c3_full_size_tf32_loss_curve_tracks_fp32_grade(b2nn):
    """Config C3 at full size for 24 momentum steps (3 epochs of 8 batches of 8192): the tf32 run's printed cost must track
    the fp32-grade (3xTF32) run's — the size-independent form of the tf32 bound stated in test_train_gpu.py — and both
    must learn (the cost of batch 0 falls from epoch to epoch)."""
    layers, B, nb, K = [784, 4096, 4096, 10], 8192, 8, 10
    rng = np.random.Generator(np.random.PCG64(7))
    m = B * nb
    x = rng.standard_normal((m, 784)).astype(np.float32)
    lab = (x @ rng.standard_normal((784, K)).astype(np.float32)).argmax(1)
    x_cm = np.empty((785, m), dtype=np.float32)          # column-major m x 785 with the ones column first
    x_cm[0] = 1.0
    x_cm[1:] = x.T
    y_cm = np.zeros((K, m), dtype=np.float32)
    y_cm[lab, np.arange(m)] = 1.0
    curves = {}
    for prec, code in (("tf32", 1), ("tf32x3", 3)):
        net = b2nn.Net()
        net.set_layers(layers, b2nn.INIT_2, 42)
        net.set_train_data(x_cm.reshape(-1), y_cm.reshape(-1), m)
        d = b2nn.make_desc(momentum=0.9, rate=0.02, iters=3, display=1, batch_num=nb, classify=True, act=b2nn.ACT_TANH,
                           out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=code)
        out = net.train(d)
        curves[prec] = (np.array(out.log_J), out.final_cost, out.error_count)
        net.close()
    j32, f32, e32 = curves["tf32x3"]
    j19, f19, e19 = curves["tf32"]
    assert len(j32) == 3 and j32[-1] < j32[0] and j19[-1] < j19[0]
    np.testing.assert_allclose(j19, j32, rtol=3e-2)
    assert abs(f19 - f32) <= 3e-2 * abs(f32)
    assert abs(e19 - e32) <= 0.01 * m


def test_c2_full_size_against_reference_live_and_oracle(b2nn, oracle):
    """BASELINE.json config C2 (bikeshare regression, full batch).  With the reference's rate 0.003 on targets ~10^2 the
    trajectory starts to oscillate around iteration 180 and is chaotic from there: any two fp32 runs (the reference's, this
    engine's with its float-atomic reductions, even two runs of the engine) end 5-25 % apart.  Parity is therefore held
    tightly over the smooth first 150 iterations; the 2000-iteration run is only required to settle in the same band."""
    b = cases.bikeshare()
    layers, m = [10, 40, 160, 1], 4337
    x = cases.zscore(b["x"])
    y = b["y"].astype(np.float32)
    x_cm, y_cm = oracle.with_bias_colmajor(x), oracle.colmajor(y)

    def run(iters, display):
        d = b2nn.make_desc(momentum=0.9, rate=0.003, iters=iters, display=display, batch_num=1, classify=False,
                           act=b2nn.ACT_SIGMOID, out=b2nn.OUT_LINEAR, cost=b2nn.COST_SQUARED,
                           precision=b2nn.PREC_AUTO)     # default mode: C2 never leaves fp32
        th0, th, out = _engine_run(b2nn, layers, x_cm, y_cm, m, d, b2nn.INIT_1, 42)
        assert out.gemm_tc_launches == 0
        ref64 = oracle.train(layers, x_cm, y_cm, m, batch_num=1, iters=iters, display=display, momentum=0.9, rate=0.003,
                             hidden_act=oracle.SIGMOID, classify=False, theta0=th0, precision="f64")
        r = None
        if oracle.ref_available():
            r = oracle.ref_train(layers, x, y, batch_num=1, iters=iters, display=display, momentum=0.9, rate=0.003,
                                 hidden_act=oracle.SIGMOID, classify=False, init_kind=1, seed=42)
            assert np.array_equal(th0, r.theta0)
        return th, out, ref64, r

    # smooth regime: tight
    th, out, ref64, r = run(150, 50)
    np.testing.assert_allclose(out.log_J, ref64.log_J, rtol=2e-3)
    assert abs(out.final_cost - ref64.final_cost) <= 2e-3 * ref64.final_cost
    assert np.abs(th - ref64.theta).max() <= 5e-3 * np.abs(ref64.theta).max()
    if r is not None:
        np.testing.assert_allclose(out.log_J, r.log_J, rtol=2e-3)
        assert abs(out.final_cost - r.final_cost) <= 2e-3 * abs(r.final_cost)
    # long run: same band
    th, out, ref64, r = run(2000, 100)
    np.testing.assert_allclose(out.log_J[:1], ref64.log_J[:1], rtol=5e-3)
    assert np.isfinite(out.final_cost) and 0.5 * ref64.final_cost <= out.final_cost <= 2.0 * ref64.final_cost
    if r is not None:
        np.testing.assert_allclose(out.log_J[:1], r.log_J[:1], rtol=5e-3)
        assert 0.5 * ref64.final_cost <= r.final_cost <= 2.0 * ref64.final_cost
