"""Parity tests proper: the CUDA engine, called through the C ABI, against
  (1) the CPU oracle on the same seeded inputs,
  (2) the UNMODIFIED reference (oracle/_ref, cuBLAS path) run live on the same GPU when its library travelled,
  (3) the committed golden vectors of that reference,
and, at BASELINE.json's full C3 size, against an fp64 restatement of one step (size-independent properties)."""
import os

import numpy as np
import pytest

import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

# Stated tolerances (max-norm relative to max|theta|, relative on J):
#   fp32 mode: contraction in fp32 FFMA — same arithmetic class as cublasSgemm default math; only summation order and
#              the libm ulps of expf/tanhf/logf differ.
#   tf32 mode: inputs of the large contractions rounded to 10 mantissa bits (tcgen05 kind::tf32), fp32 accumulate.
#              Over hundreds of steps on an ill-conditioned regression (bikeshare targets ~ 10^2) the reduced-precision
#              trajectory drifts away in theta while the cost — the quantity the user sees — stays within bound, so
#              long tf32 runs are judged on the cost curve only (LONG_TF32).
#   tf32x3   : fp32-grade on tensor cores (hi*hi + hi*lo + lo*hi per product, fp32 accumulate) — held to the fp32 bounds.
#   bf16     : tcgen05 kind::f16 on bf16 copies of the operands (8 mantissa bits, round to nearest), fp32 accumulate and fp32
#              master copies of everything: stated bound 6e-2 on the weights, 6e-2 on J.
TOL = {"fp32": dict(theta=2e-4, J=2e-4), "tf32": dict(theta=2e-2, J=3e-2), "tf32x3": dict(theta=2e-4, J=2e-4),
       "bf16": dict(theta=6e-2, J=6e-2)}
PREC = {"fp32": 0, "tf32": 1, "tf32x3": 3, "bf16": 4}
LONG_TF32 = {"reg_sigmoid_bikeshare512"}


def _desc(b2nn, c, prec, **over):
    kw = dict(momentum=c.momentum, rate=c.rate, iters=c.iters, display=c.display, batch_num=c.batch_num,
              classify=c.classify, act=b2nn.ACT_SIGMOID if c.hidden_act == cases.SIGMOID else b2nn.ACT_TANH,
              out=(b2nn.OUT_SIGMOID if c.out_act == cases.OUT_SIGMOID else b2nn.OUT_SOFTMAX) if c.classify else b2nn.OUT_LINEAR,
              cost=(b2nn.COST_NEGLNMAX if c.cost == cases.COST_NEGLNMAX else b2nn.COST_CROSSENTROPY) if c.classify else b2nn.COST_SQUARED,
              precision=PREC[prec])
    kw.update(over)
    return b2nn.make_desc(**kw)


def _oracle(O, c, theta0, precision="f32", **over):
    x_cm, y_cm = cases.oracle_inputs(c)
    kw = dict(batch_num=c.batch_num, iters=c.iters, display=c.display, momentum=c.momentum, rate=c.rate,
              hidden_act=c.hidden_act, classify=c.classify, out_act=c.out_act, cost=c.cost, precision=precision)
    kw.update(over)
    return O.train(c.layers, x_cm, y_cm, c.m, theta0=theta0, **kw)


def _engine(b2nn, c, prec, theta0=None, **over):
    net = b2nn.Net()
    net.set_layers(c.layers, c.init_kind, c.seed)
    th0 = net.get_weights()
    if theta0 is not None:
        net.set_weights(theta0)
    x_cm, y_cm = cases.oracle_inputs(c)
    net.set_train_data(x_cm, y_cm, c.m)
    out = net.train(_desc(b2nn, c, prec, **over))
    th = net.get_weights()
    return net, th0, th, out


ALL = [c.name for c in cases.small_cases()]


@pytest.mark.parametrize("name", ALL)
def test_initial_weights_bit_identical_to_curand_host(b2nn, oracle, name):
    """setLayers parity (cublasNN.cpp:314-326): same cuRAND XORWOW stream, same rescale arithmetic => same bits."""
    c = cases.case_by_name(name)
    net = b2nn.Net()
    net.set_layers(c.layers, c.init_kind, c.seed)
    assert np.array_equal(net.get_weights(), oracle.init_weights(c.layers, c.init_kind, c.seed))
    net.close()


def test_set_get_weights_roundtrip(b2nn):
    net = b2nn.Net()
    net.set_layers([7, 5, 9, 3], 1, 1)
    th = np.arange(net.num_weights(), dtype=np.float32) * 0.25 - 3
    net.set_weights(th)
    assert np.array_equal(net.get_weights(), th)
    net.close()


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3", "bf16"])
@pytest.mark.parametrize("name", ALL)
def test_train_matches_oracle(b2nn, oracle, name, prec):
    c = cases.case_by_name(name)
    net, th0, th, out = _engine(b2nn, c, prec)
    ref = _oracle(oracle, c, th0, "f32")
    ref64 = _oracle(oracle, c, th0, "f64")
    tol = TOL[prec]
    scale = np.abs(ref64.theta).max()
    err = np.abs(th - ref64.theta).max() / scale
    err_oracle32 = np.abs(ref.theta - ref64.theta).max() / scale
    if not (prec in ("tf32", "bf16") and name in LONG_TF32):
        assert err <= max(tol["theta"], 3 * err_oracle32), (name, prec, err, err_oracle32)
    assert list(out.log_iter) == list(ref.log_iter)
    np.testing.assert_allclose(out.log_J, ref64.log_J, rtol=tol["J"])
    assert abs(out.final_cost - ref64.final_cost) <= tol["J"] * abs(ref64.final_cost) + 1e-7
    if c.classify:
        assert abs(out.error_count - ref64.error_count) <= max(1.0, 0.005 * c.m)
    assert out.steps == c.iters * c.batch_num
    if prec in ("tf32", "tf32x3", "bf16") and name == "wide_mini":
        assert out.gemm_tc_launches > 0, "tensor-core modes must run the large contractions on the tcgen05 kernel"
    if prec == "fp32":
        assert out.gemm_tc_launches == 0
    net.close()


@pytest.mark.parametrize("name", ALL)
def test_train_matches_reference_live(b2nn, oracle, name):
    """Same GPU, same seed, same data: reference cuBLAS build vs this engine in fp32 mode."""
    if not oracle.ref_available():
        pytest.skip("oracle/_ref/libcublasml_ref.so did not travel")
    c = cases.case_by_name(name)
    r = oracle.ref_train(c.layers, c.x, c.y, batch_num=c.batch_num, iters=c.iters, display=c.display,
                         momentum=c.momentum, rate=c.rate, hidden_act=c.hidden_act, classify=c.classify,
                         out_act=c.out_act, cost=c.cost, init_kind=c.init_kind, seed=c.seed)
    net, th0, th, out = _engine(b2nn, c, "fp32")
    assert np.array_equal(th0, r.theta0), "initial weights must be bit-identical to the reference's"
    scale = np.abs(r.theta).max()
    assert np.abs(th - r.theta).max() <= TOL["fp32"]["theta"] * scale
    if c.compare_log:
        assert list(out.log_iter) == list(r.log_iter)
        np.testing.assert_allclose(out.log_J, r.log_J, rtol=5e-4)     # the reference prints 6 significant digits
    assert abs(out.final_cost - r.final_cost) <= 5e-4 * abs(r.final_cost) + 1e-6
    if c.classify:
        assert abs(out.error_count - r.error_count) <= max(1.0, 0.002 * c.m)
    net.close()


@pytest.mark.parametrize("name", ALL)
def test_engine_matches_golden(b2nn, name):
    path = cases.golden_path(name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not committed yet")
    g = np.load(path)
    c = cases.case_by_name(name)
    net, th0, th, out = _engine(b2nn, c, "fp32")
    assert np.array_equal(th0, g["theta0"])
    assert np.abs(th - g["theta"]).max() <= TOL["fp32"]["theta"] * np.abs(g["theta"]).max()
    if c.compare_log:
        np.testing.assert_allclose(out.log_J, g["log_J"], rtol=5e-4)
    net.close()


@pytest.mark.parametrize("name", ["clf_tanh_softmax", "reg_sigmoid_bikeshare512", "mnist_mini"])
def test_predict_and_evaluate_match_oracle(b2nn, oracle, name):
    c = cases.case_by_name(name)
    net, th0, th, out = _engine(b2nn, c, "fp32")
    x_cm, y_cm = cases.oracle_inputs(c)
    d = _desc(b2nn, c, "fp32")
    got, probs = net.predict(d, x_cm, c.m)
    want, wprobs = oracle.predict(c.layers, th, x_cm, c.m, hidden_act=c.hidden_act, classify=c.classify,
                                  out_act=c.out_act, precision="f64")
    np.testing.assert_allclose(probs, wprobs, rtol=2e-4, atol=2e-5 * max(1.0, float(np.abs(wprobs).max())))
    if c.classify:
        assert (got != want).mean() <= 0.002
        cost, errs = net.evaluate(d, x_cm, c.y[:, 0], c.m, y_is_labels=True)      # validate(): raw labels
        cost2, errs2 = net.evaluate(d, x_cm, y_cm, c.m, y_is_labels=False)
        assert abs(cost - out.final_cost) <= 1e-5 * abs(out.final_cost) + 1e-7 and errs == out.error_count
        assert abs(cost2 - cost) <= 1e-6 * abs(cost) + 1e-9 and errs2 == errs
    else:
        np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5 * max(1.0, float(np.abs(want).max())))
        cost, _ = net.evaluate(d, x_cm, y_cm, c.m)
        assert abs(cost - out.final_cost) <= 1e-5 * abs(out.final_cost)
    net.close()


def test_train_step_host_equals_resident_training(b2nn):
    """The end-to-end entry (host buffers, H2D inside the call) must walk the same trajectory as b2nn_train."""
    import ctypes
    c = cases.case_by_name("clf_tanh_softmax")
    x_cm, y_cm = cases.oracle_inputs(c)
    net, th0, th, out = _engine(b2nn, c, "fp32", skip_final_cost=True)
    net2 = b2nn.Net()
    net2.set_layers(c.layers, c.init_kind, c.seed)
    d = _desc(b2nn, c, "fp32")
    n1, K = c.layers[0] + 1, c.layers[-1]
    bs = c.m // c.batch_num
    X = x_cm.reshape(n1, c.m)
    Y = y_cm.reshape(K, c.m)
    Js = []
    for it in range(c.iters):
        for b in range(c.batch_num):
            lo, hi = b * bs, (c.m if b == c.batch_num - 1 else (b + 1) * bs)
            xb = np.ascontiguousarray(X[:, lo:hi]).reshape(-1)
            yb = np.ascontiguousarray(Y[:, lo:hi]).reshape(-1)
            J = net2.train_step_host(d, ctypes.c_void_p(xb.ctypes.data), ctypes.c_void_p(yb.ctypes.data), hi - lo)
            if (it + 1) % c.display == 0 and b == 0:
                Js.append(J)
    th2 = net2.get_weights()
    assert np.abs(th2 - th).max() <= 1e-6 * np.abs(th).max()
    np.testing.assert_allclose(Js, out.log_J, rtol=1e-5)
    net.close(); net2.close()


@pytest.mark.parametrize("name", ["clf_tanh_softmax", "reg_tanh_ragged", "wide_mini"])
def test_streaming_host_trainer_equals_resident_training(b2nn, name):
    """b2nn_train_host (data stays on the host, double-buffered H2D) must walk the same trajectory as b2nn_train and
    report the cost of every step."""
    import ctypes
    c = cases.case_by_name(name)
    x_cm, y_cm = cases.oracle_inputs(c)
    net, th0, th, out = _engine(b2nn, c, "fp32", skip_final_cost=True, display=1)
    net2 = b2nn.Net()
    net2.set_layers(c.layers, c.init_kind, c.seed)
    d = _desc(b2nn, c, "fp32", display=1)
    J, r = net2.train_host(d, ctypes.c_void_p(x_cm.ctypes.data), ctypes.c_void_p(y_cm.ctypes.data), c.m)
    th2 = net2.get_weights()
    assert r.steps == c.iters * c.batch_num and J.shape == (r.steps,)
    assert np.abs(th2 - th).max() <= 2e-6 * np.abs(th).max()
    np.testing.assert_allclose(J[::c.batch_num], out.log_J, rtol=1e-5)       # batch 0 of every epoch = the printed cost
    assert np.all(np.isfinite(J)) and J[-1] < J[0]
    net.close(); net2.close()


def _torch_step_fp64(layers, W, X, Y, rate, hidden):
    """fp64 restatement of one classify step (softmax / cross-entropy, mu = 0) on the GPU: returns new weights.
    W[j]: (l_j+1, l_{j+1}) reference layout (bias row 0); X: (B, l_0); Y: (B, K) one-hot."""
    B = X.shape[0]
    ones = torch.ones(B, 1, dtype=torch.float64, device=X.device)
    acts = [torch.cat([ones, X], 1)]
    for j in range(len(W) - 1):
        z = acts[-1] @ W[j]
        a = torch.tanh(z) if hidden == "tanh" else torch.sigmoid(z)
        acts.append(torch.cat([ones, a], 1))
    z = acts[-1] @ W[-1]
    h = torch.softmax(z, 1)
    J = -(Y * torch.log(h)).sum() / B
    delta = h - Y
    new = []
    for j in range(len(W) - 1, -1, -1):
        G = acts[j].t() @ delta
        if j > 0:
            a = acts[j][:, 1:]
            d = (delta @ W[j].t())[:, 1:]
            delta = d * ((1 - a * a) if hidden == "tanh" else a * (1 - a))
        new.append(W[j] - (rate / B) * G)
    return new[::-1], float(J)


@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_c3_full_size_one_step_against_fp64(b2nn, oracle, prec):
    """BASELINE.json config C3 at full size (784-4096-4096-10, batch 8192): one momentum-free step on the engine vs an
    fp64 restatement of the same step; also the printed cost.  The CPU oracle is too slow at this size, so the
    check is the step's defining property: theta_1 = theta_0 - rate/B * dJ/dtheta."""
    layers, B, K = [784, 4096, 4096, 10], 8192, 10
    rng = np.random.Generator(np.random.PCG64(1))
    X = rng.standard_normal((B, 784)).astype(np.float32)
    teacher = rng.standard_normal((784, K)).astype(np.float32)
    lab = (X @ teacher).argmax(1)
    net = b2nn.Net()
    net.set_layers(layers, b2nn.INIT_2, 42)
    th0 = net.get_weights()
    x_cm = oracle.with_bias_colmajor(X)
    y_cm = oracle.one_hot_colmajor(lab, K)
    net.set_train_data(x_cm, y_cm, B)
    rate = 0.5
    d = b2nn.make_desc(momentum=0.0, rate=rate, iters=1, display=1, batch_num=1, classify=True, act=b2nn.ACT_TANH,
                       out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY,
                       precision=PREC[prec], skip_final_cost=True)
    out = net.train(d)
    th1 = net.get_weights()
    rel = 2e-2 if prec == "tf32" else 5e-4
    assert out.gemm_tc_launches >= 5
    W = []
    for pos, r, cdim in oracle.theta_offsets(layers):
        W.append(torch.tensor(th0[pos:pos + r * cdim].reshape(cdim, r).T.copy(), dtype=torch.float64, device="cuda"))
    Wn, J = _torch_step_fp64(layers, W, torch.tensor(X, dtype=torch.float64, device="cuda"),
                             torch.tensor(np.eye(K)[lab], dtype=torch.float64, device="cuda"), rate, "tanh")
    assert abs(out.log_J[0] - J) <= (2e-3 if prec == "tf32" else 2e-5) * abs(J)
    for (pos, r, cdim), w0, wn in zip(oracle.theta_offsets(layers), W, Wn):
        got = torch.tensor(th1[pos:pos + r * cdim].reshape(cdim, r).T.copy(), dtype=torch.float64, device="cuda")
        step_ref = wn - w0
        step_got = got - w0
        denom = float(step_ref.abs().max())
        assert denom > 0
        # the UPDATE (not just theta) must match: relative to the largest update entry of the layer
        assert float((step_got - step_ref).abs().max()) <= rel * denom + 2e-7, (pos, denom)   # + fp32 ulp of theta itself
        # and theta itself to fp32 resolution of the update
        assert float((got - wn).abs().max()) <= rel * denom + 2e-7
    net.close()


def test_error_paths_fail_loudly(b2nn):
    """Call-order and argument errors are undefined behaviour in the reference (cublasNN.cpp:137, 657, main.cpp:125);
    here every one of them returns a status and a message, and the context stays usable."""
    net = b2nn.Net()
    with pytest.raises(b2nn.B2nnError):
        net.set_layers([4, 3], 1, 1)                      # no hidden layer (reference dereferences aPos[-1], :657)
    with pytest.raises(b2nn.B2nnError):
        net.layers = [4, 3, 2]
        net.set_train_data(np.zeros(5 * 5, np.float32), np.zeros(5 * 2, np.float32), 5)   # set_layers never succeeded
    net.set_layers([4, 3, 2], 1, 1)
    with pytest.raises(b2nn.B2nnError):
        net.train(b2nn.make_desc(iters=1, batch_num=1))   # no data yet
    x = np.ones((5, 6), np.float32).reshape(-1)
    y = np.zeros((2, 6), np.float32).reshape(-1)
    net.set_train_data(x, y, 6)
    with pytest.raises(b2nn.B2nnError):
        net.train(b2nn.make_desc(iters=1, batch_num=7))   # more batches than rows
    with pytest.raises(b2nn.B2nnError):
        net.train(b2nn.make_desc(iters=1, batch_num=0))
    with pytest.raises(b2nn.B2nnError):
        net.set_weights(np.zeros(3, np.float32))          # wrong count
    with pytest.raises(b2nn.B2nnError):
        net.train(b2nn.make_desc(iters=1, batch_num=1, act=b2nn.ACT_CUSTOM))   # custom activation without pointers
    out = net.train(b2nn.make_desc(iters=2, batch_num=2, display=1))            # still usable
    assert out.steps == 4 and np.all(np.isfinite(out.log_J))
    net.close()


def test_single_row_batches_and_tiny_shapes(b2nn, oracle):
    """Stochastic mode (batchNum = m, one row per step) and 1-unit layers: the smallest shapes every kernel must take."""
    layers, m = [1, 1, 1], 5
    rng = np.random.default_rng(3)
    x = rng.standard_normal((m, 1)).astype(np.float32)
    yv = (0.5 * x + 0.1).astype(np.float32)
    x_cm, y_cm = oracle.with_bias_colmajor(x), oracle.colmajor(yv)
    net = b2nn.Net()
    net.set_layers(layers, 1, 9)
    th0 = net.get_weights()
    net.set_train_data(x_cm, y_cm, m)
    for prec in ("fp32", "tf32", "tf32x3"):
        net.set_weights(th0)
        d = b2nn.make_desc(momentum=0.5, rate=0.1, iters=7, display=1, batch_num=m, classify=False, act=b2nn.ACT_TANH,
                           out=b2nn.OUT_LINEAR, cost=b2nn.COST_SQUARED, precision=PREC[prec])
        out = net.train(d)
        ref = oracle.train(layers, x_cm, y_cm, m, batch_num=m, iters=7, display=1, momentum=0.5, rate=0.1,
                           hidden_act=oracle.TANH, classify=False, theta0=th0, precision="f64")
        assert np.abs(net.get_weights() - ref.theta).max() <= 2e-5
        np.testing.assert_allclose(out.log_J, ref.log_J, rtol=1e-4, atol=1e-7)
        assert abs(out.final_cost - ref.final_cost) <= 1e-4 * ref.final_cost + 1e-7
    net.close()


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3"])
@pytest.mark.parametrize("act", ["tanh", "sigmoid"])
def test_wide_hidden_narrow_output_ragged_batches(b2nn, oracle, act, prec):
    """A 1100-unit hidden layer in front of 7 outputs takes the fused delta-backprop + gradient kernel (skinny.cu); the
    ragged batches (111 / 111 / 115 rows) make it run with deltas of a longer previous batch left behind the batch end."""
    layers, m, K = [40, 1100, 7], 337, 7
    rng = np.random.default_rng(11)
    x = rng.standard_normal((m, 40)).astype(np.float32)
    lab = (x @ rng.standard_normal((40, K)).astype(np.float32)).argmax(1)
    x_cm, y_cm = oracle.with_bias_colmajor(x), oracle.one_hot_colmajor(lab, K)
    net = b2nn.Net()
    net.set_layers(layers, b2nn.INIT_2, 5)
    th0 = net.get_weights()
    net.set_train_data(x_cm, y_cm, m)
    kw = dict(momentum=0.9, rate=0.05, iters=4, display=1, batch_num=3, classify=True)
    d = b2nn.make_desc(act=b2nn.ACT_TANH if act == "tanh" else b2nn.ACT_SIGMOID, out=b2nn.OUT_SOFTMAX,
                       cost=b2nn.COST_CROSSENTROPY, precision=PREC[prec], **kw)
    out = net.train(d)
    ref = oracle.train(layers, x_cm, y_cm, m, hidden_act=oracle.TANH if act == "tanh" else oracle.SIGMOID,
                       out_act=oracle.OUT_SOFTMAX, cost=oracle.COST_CROSSENTROPY, theta0=th0, precision="f64", **kw)
    tol = TOL[prec]
    err = np.abs(net.get_weights() - ref.theta).max() / np.abs(ref.theta).max()
    assert err <= tol["theta"], (act, prec, err)
    np.testing.assert_allclose(out.log_J, ref.log_J, rtol=tol["J"])
    net.close()


@pytest.mark.parametrize("classify", [True, False])
def test_tiny_net_whole_step_kernel_many_sample_groups(b2nn, oracle, classify):
    """A net whose weights fit shared memory runs each step as one kernel (micro.cu).  10 007 rows per batch are 313
    groups of 32 samples over at most 148 blocks, so every block walks several groups (and the last one is ragged)."""
    layers, m = ([5, 8, 6, 3] if classify else [5, 8, 6, 2]), 20014
    rng = np.random.default_rng(21)
    x = rng.standard_normal((m, 5)).astype(np.float32)
    x_cm = oracle.with_bias_colmajor(x)
    if classify:
        lab = (x @ rng.standard_normal((5, 3)).astype(np.float32)).argmax(1)
        y_cm = oracle.one_hot_colmajor(lab, 3)
    else:
        y_cm = oracle.colmajor(np.tanh(x @ rng.standard_normal((5, 2)).astype(np.float32)).astype(np.float32))
    net = b2nn.Net()
    net.set_layers(layers, b2nn.INIT_2, 3)
    th0 = net.get_weights()
    net.set_train_data(x_cm, y_cm, m)
    kw = dict(momentum=0.9, rate=0.05, iters=6, display=1, batch_num=2, classify=classify)
    d = b2nn.make_desc(act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX if classify else b2nn.OUT_LINEAR,
                       cost=b2nn.COST_CROSSENTROPY if classify else b2nn.COST_SQUARED, precision=b2nn.PREC_AUTO, **kw)
    out = net.train(d)
    assert out.kernel_launches == 2 * out.steps, "tiny nets take two launches per step: the whole-step kernel and the update"
    ref = oracle.train(layers, x_cm, y_cm, m, hidden_act=oracle.TANH, theta0=th0, precision="f64", **kw)
    assert np.abs(net.get_weights() - ref.theta).max() <= 2e-4 * np.abs(ref.theta).max()
    np.testing.assert_allclose(out.log_J, ref.log_J, rtol=2e-4)
    assert abs(out.final_cost - ref.final_cost) <= 2e-4 * abs(ref.final_cost)
    net.close()


def _addr(b2nn, mangled):
    import ctypes
    return ctypes.cast(getattr(b2nn.lib(), mangled), ctypes.c_void_p).value


@pytest.mark.parametrize("which", ["hidden", "output", "cost", "all"])
@pytest.mark.parametrize("name", ["clf_tanh_softmax", "clf_sigmoid_sigout"])
def test_function_pointer_path_matches_fused_builtins(b2nn, name, which):
    """The reference's extension point: user functions passed by pointer (cublasNN.h:45-50).  Feeding the library's own
    exported ops back in as 'unknown' pointers forces the generic path (materialised z, calls on the legacy default
    stream, like cublasNN.cpp:606, 613, 648, 673); it must reproduce the fused built-in path."""
    c = cases.case_by_name(name)
    tanh_case = c.hidden_act == cases.TANH
    act = "_Z7tanhGPUPKfPfi" if tanh_case else "_Z10sigmoidGPUPKfPfi"
    dact = "_Z9sechSqGPUPKfPfi" if tanh_case else "_Z14sigmoidGradGPUPKfPfi"
    outf = "_Z10softmaxGPUPKfPfii" if c.out_act == cases.OUT_SOFTMAX else "_Z16sigmoidOutputGPUPKfPfii"
    costf = "_Z19crossEntropyCostGPUPfS_S_i" if c.cost == cases.COST_CROSSENTROPY else "_Z15negLnMaxCostGPUPfS_S_i"
    net, th0, th, out = _engine(b2nn, c, "fp32")
    x_cm, y_cm = cases.oracle_inputs(c)
    net2 = b2nn.Net()
    net2.set_layers(c.layers, c.init_kind, c.seed)
    net2.set_train_data(x_cm, y_cm, c.m)
    d = _desc(b2nn, c, "fp32")
    if which in ("hidden", "all"):
        d.act_kind = b2nn.ACT_CUSTOM
        d.act_fn, d.act_grad_fn = _addr(b2nn, act), _addr(b2nn, dact)
    if which in ("output", "all"):
        d.out_kind = b2nn.OUT_CUSTOM
        d.out_fn = _addr(b2nn, outf)
    if which in ("cost", "all"):
        d.cost_kind = b2nn.COST_CUSTOM
        d.cost_fn = _addr(b2nn, costf)
    out2 = net2.train(d)
    th2 = net2.get_weights()
    assert np.abs(th2 - th).max() <= 2e-5 * np.abs(th).max(), which
    np.testing.assert_allclose(out2.log_J, out.log_J, rtol=2e-5)
    assert abs(out2.final_cost - out.final_cost) <= 2e-5 * abs(out.final_cost) + 1e-8
    assert out2.error_count == out.error_count
    got, probs = net2.predict(d, x_cm, c.m)
    want, wprobs = net.predict(_desc(b2nn, c, "fp32"), x_cm, c.m)
    np.testing.assert_allclose(probs, wprobs, rtol=1e-4, atol=1e-6)
    assert (got != want).mean() <= 0.005
    net.close(); net2.close()


@pytest.fixture(scope="module")
def c4_mini_oracle(oracle):
    """fp64 oracle run of the nine-weight-layer C4 miniature (about 20 s of CPU), shared by the three precision cases."""
    c = cases.case_by_name("c4_mini")
    th0 = oracle.init_weights(c.layers, c.init_kind, c.seed)
    return c, th0, _oracle(oracle, c, th0, "f64")


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3", "bf16"])
def test_c4_mini_nine_layers_match_oracle(b2nn, c4_mini_oracle, prec):
    """Config C4's depth (8 hidden layers, cublasNN.cpp:642-685 walks them in a loop) at a size the oracle finishes:
    every layer of the forward / delta-backprop / gradient chain against fp64."""
    c, th0, ref64 = c4_mini_oracle
    net, th0_engine, th, out = _engine(b2nn, c, prec)
    assert np.array_equal(th0_engine, th0)
    tol = TOL[prec]
    assert np.abs(th - ref64.theta).max() <= tol["theta"] * np.abs(ref64.theta).max()
    np.testing.assert_allclose(out.log_J, ref64.log_J, rtol=tol["J"])
    assert abs(out.final_cost - ref64.final_cost) <= tol["J"] * abs(ref64.final_cost)
    assert abs(out.error_count - ref64.error_count) <= max(1.0, 0.005 * c.m)
    if prec != "fp32":
        assert out.gemm_tc_launches >= 2 * 8 * out.steps, "the wide layers must run on the tcgen05 kernel"
    net.close()


@pytest.mark.parametrize("prec", ["tf32", "tf32x3"])
def test_one_kernel_step_for_one_hidden_layer_nets(b2nn, oracle, monkeypatch, prec):
    """mlp3.cu (opt-in, B2NN_MLP3=1): the whole step of a one-hidden-layer net in one kernel, warp-level tf32 MMA with the
    fp32-grade split done in registers.  Same bounds as the default path; inference (fp32-grade) goes through its forward half."""
    monkeypatch.setenv("B2NN_MLP3", "1")
    c = cases.case_by_name("mnist_mini")
    net, th0, th, out = _engine(b2nn, c, prec)
    ref64 = _oracle(oracle, c, th0, "f64")
    tol = TOL[prec]
    assert out.kernel_launches == 2 * out.steps, "one step kernel + one reduce/update kernel per step"
    assert np.abs(th - ref64.theta).max() <= tol["theta"] * np.abs(ref64.theta).max()
    np.testing.assert_allclose(out.log_J, ref64.log_J, rtol=tol["J"])
    assert abs(out.final_cost - ref64.final_cost) <= tol["J"] * abs(ref64.final_cost) + 1e-7
    assert abs(out.error_count - ref64.error_count) <= max(1.0, 0.005 * c.m)
    x_cm, _ = cases.oracle_inputs(c)
    got, probs = net.predict(_desc(b2nn, c, prec), x_cm, c.m)
    want, wprobs = oracle.predict(c.layers, th, x_cm, c.m, hidden_act=c.hidden_act, classify=True, out_act=c.out_act, precision="f64")
    assert np.abs(probs - wprobs).max() <= (5e-3 if prec == "tf32" else 5e-5)
    assert (got != want).mean() <= 0.01
    net.close()
