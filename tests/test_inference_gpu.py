"""Inference / evaluation parity (config C5: predict / validate, cublasNN.cpp:779-977) and the data-staging entry points
around the hot path: the chunked, double-buffered b2nn_predict, multi-chunk evaluate, device-side normalise / bias
(cublasNN.cpp:191-268, 330-346), checkpoint round trips, and regressions for the round-1 advisor findings."""
import ctypes
import os

import numpy as np
import pytest

import cases

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

PREC = {"fp32": 0, "tf32": 1, "tf32x3": 3}
# max |p - p64| on the class probabilities of the 784-50-10 net with 3x Glorot weights (|z| up to ~15): fp32-grade modes
# differ from the fp64 oracle by fp32 rounding of 785-long sums only (observed <= 2.5e-5 over 655 370 probabilities); tf32
# truncates both operands to 10 mantissa bits — a hundred times looser.
PROB_ATOL = {"fp32": 5e-5, "tf32x3": 5e-5, "tf32": 5e-3}
MNIST_NET = [784, 50, 10]


def _desc(b2nn, prec, classify=True, act=None, out=None):
    return b2nn.make_desc(classify=classify, act=b2nn.ACT_TANH if act is None else act,
                          out=b2nn.OUT_SOFTMAX if out is None else out, cost=b2nn.COST_CROSSENTROPY,
                          precision=PREC[prec])


def _mnist_net(b2nn):
    net = b2nn.Net()
    net.set_layers(MNIST_NET, b2nn.INIT_2, 42)
    th = net.get_weights() * np.float32(3.0)        # sharper outputs than Glorot's: argmax ties become rare
    net.set_weights(th)
    return net, th


def _rows_x(rows, n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.standard_normal((rows, n), dtype=np.float32)


@pytest.mark.parametrize("prec", ["fp32", "tf32", "tf32x3"])
@pytest.mark.parametrize("rows", [1, 31, 33, 4200, 65537])
def test_predict_sweep_matches_oracle(b2nn, oracle, prec, rows):
    """C5 on the MNIST net: class ids and probabilities of b2nn_predict against the fp64 oracle for batch sizes on both
    sides of every tile boundary (1, not a multiple of 32, one past a multiple of 2^16)."""
    net, th = _mnist_net(b2nn)
    x = _rows_x(rows, 784, seed=rows)
    x_cm = oracle.with_bias_colmajor(x)
    got, probs = net.predict(_desc(b2nn, prec), x_cm, rows)
    want, wprobs = oracle.predict(MNIST_NET, th, x_cm, rows, hidden_act=oracle.TANH, classify=True,
                                  out_act=oracle.OUT_SOFTMAX, precision="f64")
    assert np.abs(probs - wprobs).max() <= PROB_ATOL[prec], (prec, rows, float(np.abs(probs - wprobs).max()))
    # a class id may only differ where the two best probabilities are closer than the stated bound
    bad = np.nonzero(got != want)[0]
    P = wprobs.reshape(10, rows)
    for i in bad:
        top = np.sort(P[:, i])[-2:]
        assert top[1] - top[0] <= 2 * PROB_ATOL[prec], (prec, rows, int(i), top)
    net.close()


@pytest.mark.parametrize("prec", ["tf32", "fp32"])
def test_predict_quarter_million_rows(b2nn, oracle, prec):
    """2^18 rows (the upper half of C5's sweep): several pipeline chunks; checked against the fp64 oracle on a strided
    sample of the rows and, for the rest, through the chunk-independence property below."""
    rows = 1 << 18
    net, th = _mnist_net(b2nn)
    x = _rows_x(rows, 784, seed=5)
    x_cm = oracle.with_bias_colmajor(x)
    got, probs = net.predict(_desc(b2nn, prec), x_cm, rows)
    idx = np.arange(0, rows, 61)
    xs_cm = oracle.with_bias_colmajor(x[idx])
    want, wprobs = oracle.predict(MNIST_NET, th, xs_cm, len(idx), hidden_act=oracle.TANH, classify=True,
                                  out_act=oracle.OUT_SOFTMAX, precision="f64")
    P = probs.reshape(10, rows)[:, idx]
    assert np.abs(P - wprobs.reshape(10, len(idx))).max() <= PROB_ATOL[prec]
    assert (got[idx] != want).mean() <= (0.02 if prec == "tf32" else 0.001)
    assert np.all((got >= 0) & (got <= 9)) and np.allclose(probs.reshape(10, rows).sum(0), 1.0, atol=1e-4)
    net.close()


@pytest.mark.parametrize("pinned", [False, True])
def test_predict_chunk_pipeline_is_chunk_independent(b2nn, oracle, monkeypatch, pinned):
    """Forcing 1024-row chunks (5 chunks through the two staging buffers, the last one ragged) must give bit-identical
    results to the single-chunk call in fp32 mode, from pageable and from pinned host memory."""
    rows = 4200 + 17
    net, th = _mnist_net(b2nn)
    x_cm = oracle.with_bias_colmajor(_rows_x(rows, 784, seed=11))
    d = _desc(b2nn, "fp32")
    whole, wprobs = net.predict(d, x_cm, rows)
    monkeypatch.setenv("B2NN_PREDICT_CHUNK", "1024")
    net2 = b2nn.Net()
    net2.set_layers(MNIST_NET, b2nn.INIT_2, 42)
    net2.set_weights(th)
    if pinned:
        xh = torch.from_numpy(x_cm).pin_memory()
        out = torch.empty(rows, dtype=torch.float32).pin_memory()
        pr = torch.empty(rows * 10, dtype=torch.float32).pin_memory()
        for _ in range(2):      # second call reuses every buffer and event
            net2.predict_ptr(d, ctypes.c_void_p(xh.data_ptr()), rows, ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(pr.data_ptr()))
        got, probs = out.numpy(), pr.numpy()
    else:
        got, probs = net2.predict(d, x_cm, rows)
        got, probs = net2.predict(d, x_cm, rows)
    assert np.array_equal(got, whole) and np.array_equal(probs, wprobs)
    net.close(); net2.close()


def test_predict_regression_outputs_chunked(b2nn, oracle, monkeypatch):
    """predictFuncApprox (linear outputs, K = 3 columns) through the chunk pipeline."""
    c = cases.case_by_name("reg_tanh_ragged")
    net = b2nn.Net()
    net.set_layers(c.layers, c.init_kind, c.seed)
    th = net.get_weights()
    x_cm, _ = cases.oracle_inputs(c)
    monkeypatch.setenv("B2NN_PREDICT_CHUNK", "64")
    d = _desc(b2nn, "fp32", classify=False)
    got, _ = net.predict(d, x_cm, c.m)
    want, _ = oracle.predict(c.layers, th, x_cm, c.m, hidden_act=oracle.TANH, classify=False, precision="f64")
    np.testing.assert_allclose(got, want, rtol=2e-4, atol=2e-5)
    net.close()


def test_evaluate_two_chunks_on_a_wide_net(b2nn, oracle):
    """calcFinalCost / validate over more rows than one workspace chunk (evaluate_device's chunk loop): the cost and the
    error count of 66 000 rows must equal the row-weighted combination of the two halves evaluated on their own."""
    layers, rows = [32, 2048, 2048, 10], 66000
    net = b2nn.Net()
    net.set_layers(layers, b2nn.INIT_2, 3)
    x = _rows_x(rows, 32, seed=2)
    lab = np.random.Generator(np.random.PCG64(4)).integers(0, 10, rows)
    d = _desc(b2nn, "fp32")
    cost, errs = net.evaluate(d, oracle.with_bias_colmajor(x), lab.astype(np.float32), rows, y_is_labels=True)
    h = rows // 2
    c0, e0 = net.evaluate(d, oracle.with_bias_colmajor(x[:h]), lab[:h].astype(np.float32), h, y_is_labels=True)
    c1, e1 = net.evaluate(d, oracle.with_bias_colmajor(x[h:]), lab[h:].astype(np.float32), rows - h, y_is_labels=True)
    assert errs == e0 + e1
    assert abs(cost - (c0 * h + c1 * (rows - h)) / rows) <= 1e-6 * cost
    # and a sample of it against the fp64 oracle
    th = net.get_weights()
    got, _ = net.predict(d, oracle.with_bias_colmajor(x[:512]), 512)
    want, _ = oracle.predict(layers, th, oracle.with_bias_colmajor(x[:512]), 512, hidden_act=oracle.TANH, classify=True,
                             out_act=oracle.OUT_SOFTMAX, precision="f64")
    assert (got != want).mean() <= 0.01
    net.close()


def test_device_side_normalise_and_bias_match_host_path(b2nn, oracle):
    """b2nn_set_train_data_raw (normaliseData + addBiasData + copyDataGPU on the device) against the host z-score:
    same statistics (double vs the reference's float accumulation: 1e-6), same training trajectory."""
    c = cases.case_by_name("clf_tanh_softmax")
    rng = np.random.Generator(np.random.PCG64(21))
    raw = (c.x * rng.uniform(0.5, 20.0, c.x.shape[1]).astype(np.float32) + rng.uniform(-5, 5, c.x.shape[1]).astype(np.float32))
    raw[:, 3] = 7.0                                               # a constant feature normalises to 0 (cublasNN.cpp:207)
    xn = cases.zscore(raw)
    K = c.layers[-1]
    y_cm = oracle.one_hot_colmajor(c.y[:, 0].astype(np.int64), K)
    d = b2nn.make_desc(momentum=c.momentum, rate=c.rate, iters=c.iters, display=c.display, batch_num=c.batch_num,
                       classify=True, act=b2nn.ACT_TANH, precision=0)
    a = b2nn.Net(); a.set_layers(c.layers, c.init_kind, c.seed)
    a.set_train_data(oracle.with_bias_colmajor(xn), y_cm, c.m)
    oa = a.train(d)
    b = b2nn.Net(); b.set_layers(c.layers, c.init_kind, c.seed)
    mean, sd = b.set_train_data_raw(oracle.colmajor(raw), y_cm, c.m, normalise=True)
    np.testing.assert_allclose(mean, raw.astype(np.float64).mean(0), rtol=2e-6, atol=1e-6)
    np.testing.assert_allclose(sd, raw.astype(np.float64).std(0, ddof=1), rtol=2e-6, atol=1e-6)
    ob = b.train(d)
    assert np.abs(a.get_weights() - b.get_weights()).max() <= 2e-5 * np.abs(a.get_weights()).max()
    np.testing.assert_allclose(ob.log_J, oa.log_J, rtol=2e-5)
    assert ob.error_count == oa.error_count
    # validate / predict on raw rows reuse the stored statistics
    ca, ea = a.evaluate(d, oracle.with_bias_colmajor(xn), c.y[:, 0], c.m, y_is_labels=True)
    cb, eb = b.evaluate_raw(d, oracle.colmajor(raw), c.y[:, 0], c.m, y_is_labels=True, normalise=True)
    assert abs(ca - cb) <= 5e-5 * abs(ca) and ea == eb
    pa, _ = a.predict(d, oracle.with_bias_colmajor(xn), c.m)
    pb, _ = b.predict_raw(d, oracle.colmajor(raw), c.m, normalise=True)
    assert (pa != pb).mean() <= 0.005
    a.close(); b.close()


def test_checkpoint_round_trip(b2nn, oracle, tmp_path):
    """save -> load into a fresh context (no set_layers): identical weights, velocity, statistics and predictions; a
    corrupted file is rejected."""
    c = cases.case_by_name("clf_tanh_softmax")
    x_cm, y_cm = cases.oracle_inputs(c)
    net = b2nn.Net(); net.set_layers(c.layers, c.init_kind, c.seed)
    raw = oracle.colmajor(c.x)
    net.set_train_data_raw(raw, y_cm, c.m, normalise=True)
    d = b2nn.make_desc(momentum=0.9, rate=0.05, iters=3, display=1, batch_num=2, classify=True, act=b2nn.ACT_TANH, precision=0)
    net.train(d)
    path = str(tmp_path / "net.b2nn")
    net.save_checkpoint(path)
    other = b2nn.Net()
    other.load_checkpoint(path, layers=c.layers)
    assert np.array_equal(other.get_weights(), net.get_weights())
    m0, s0 = net.get_normalisation(); m1, s1 = other.get_normalisation()
    assert np.array_equal(m0, m1) and np.array_equal(s0, s1)
    p0, q0 = net.predict_raw(d, raw, c.m); p1, q1 = other.predict_raw(d, raw, c.m)
    assert np.array_equal(p0, p1) and np.array_equal(q0, q1)
    # velocity travels too: one more host step from both contexts lands on the same weights
    n1, K = c.layers[0] + 1, c.layers[-1]
    xb = np.ascontiguousarray(x_cm.reshape(n1, c.m)[:, :64]).reshape(-1)
    yb = np.ascontiguousarray(y_cm.reshape(K, c.m)[:, :64]).reshape(-1)
    for n in (net, other):
        n.train_step_host(d, ctypes.c_void_p(xb.ctypes.data), ctypes.c_void_p(yb.ctypes.data), 64)
    assert np.array_equal(other.get_weights(), net.get_weights())
    blob = bytearray(open(path, "rb").read())
    blob[len(blob) // 2] ^= 0x40
    bad = str(tmp_path / "bad.b2nn")
    open(bad, "wb").write(bytes(blob))
    with pytest.raises(b2nn.B2nnError):
        b2nn.Net().load_checkpoint(bad)
    # a truncated file, and a header that promises far more weights than the file holds (nothing may be sized by it)
    good = open(path, "rb").read()
    open(bad, "wb").write(good[:len(good) - 100])
    with pytest.raises(b2nn.B2nnError):
        b2nn.Net().load_checkpoint(bad)
    nl = int(np.frombuffer(good[8:12], dtype=np.int32)[0])
    lie = bytearray(good)
    lie[12:16] = np.int32(1 << 30).tobytes()                    # first layer width
    lie[12 + 4 * nl + 4:12 + 4 * nl + 8] = np.int32(0).tobytes()      # no normalisation statistics: the header stays consistent
    off_nw = 12 + 4 * nl + 8
    lie[off_nw:off_nw + 8] = np.int64(((1 << 30) + 1) * c.layers[1] + sum((a + 1) * b for a, b in zip(c.layers[1:-1], c.layers[2:]))).tobytes()
    open(bad, "wb").write(bytes(lie))
    with pytest.raises(b2nn.B2nnError):
        b2nn.Net().load_checkpoint(bad)
    net.close(); other.close()


def _addr(b2nn, symbol):
    return ctypes.cast(getattr(b2nn.lib(), symbol), ctypes.c_void_p).value


def test_forward_only_entry_points_accept_custom_activation_without_derivative(b2nn, oracle):
    """validate* / predict* take activationHidden only (cublasNN.h:51-69): a user pointer without its derivative must
    work for evaluate / predict and still be rejected by train."""
    c = cases.case_by_name("clf_tanh_softmax")
    x_cm, y_cm = cases.oracle_inputs(c)
    net = b2nn.Net(); net.set_layers(c.layers, c.init_kind, c.seed)
    net.set_train_data(x_cm, y_cm, c.m)
    builtin = b2nn.make_desc(classify=True, act=b2nn.ACT_TANH, precision=0)
    want, wprobs = net.predict(builtin, x_cm, c.m)
    wc, we = net.evaluate(builtin, x_cm, y_cm, c.m)
    d = b2nn.make_desc(classify=True, act=b2nn.ACT_CUSTOM, precision=0)
    d.act_fn = _addr(b2nn, "_Z7tanhGPUPKfPfi")
    got, probs = net.predict(d, x_cm, c.m)
    gc, ge = net.evaluate(d, x_cm, y_cm, c.m)
    np.testing.assert_allclose(probs, wprobs, rtol=1e-4, atol=1e-6)
    assert (got != want).mean() <= 0.005 and abs(gc - wc) <= 1e-4 * abs(wc) and ge == we
    with pytest.raises(b2nn.B2nnError):
        net.train(d)
    net.close()


def test_set_layers_again_resizes_host_staging(b2nn, oracle):
    """A second set_layers with wider input / output layers must not reuse staging buffers sized for the old widths
    (b2nn_train_step_host, b2nn_train_host, b2nn_predict)."""
    net = b2nn.Net()
    d = b2nn.make_desc(momentum=0.9, rate=0.05, iters=1, display=1, batch_num=1, classify=True, act=b2nn.ACT_TANH, precision=0)
    for layers, seed in (([6, 8, 3], 1), ([40, 16, 7], 2), ([5, 8, 2], 3)):
        rows = 96
        net.set_layers(layers, b2nn.INIT_2, seed)
        th0 = net.get_weights()
        x = _rows_x(rows, layers[0], seed)
        lab = np.random.Generator(np.random.PCG64(seed)).integers(0, layers[-1], rows)
        x_cm, y_cm = oracle.with_bias_colmajor(x), oracle.one_hot_colmajor(lab, layers[-1])
        J = net.train_step_host(d, ctypes.c_void_p(x_cm.ctypes.data), ctypes.c_void_p(y_cm.ctypes.data), rows)
        ref = oracle.train(layers, x_cm, y_cm, rows, batch_num=1, iters=1, display=1, momentum=0.9, rate=0.05,
                           hidden_act=oracle.TANH, classify=True, out_act=oracle.OUT_SOFTMAX, cost=oracle.COST_CROSSENTROPY,
                           theta0=th0, precision="f64")
        assert abs(J - ref.log_J[0]) <= 2e-4 * abs(ref.log_J[0])
        assert np.abs(net.get_weights() - ref.theta).max() <= 2e-4 * np.abs(ref.theta).max()
        Js, _ = net.train_host(d, ctypes.c_void_p(x_cm.ctypes.data), ctypes.c_void_p(y_cm.ctypes.data), rows)
        assert np.isfinite(Js).all()
        got, _ = net.predict(d, x_cm, rows)
        want, _ = oracle.predict(layers, net.get_weights(), x_cm, rows, hidden_act=oracle.TANH, classify=True,
                                 out_act=oracle.OUT_SOFTMAX, precision="f64")
        assert (got != want).mean() <= 0.02
    net.close()


def test_default_precision_is_fp32_grade(b2nn, oracle, monkeypatch):
    """B2NN_PREC_AUTO (what the drop-in facade passes) must hold the fp32 tolerance on a net whose contractions are large
    enough for the tensor cores; plain tf32 is an explicit opt-in."""
    monkeypatch.delenv("B2NN_PRECISION", raising=False)
    c = cases.case_by_name("wide_mini")
    x_cm, y_cm = cases.oracle_inputs(c)
    net = b2nn.Net(); net.set_layers(c.layers, c.init_kind, c.seed)
    th0 = net.get_weights()
    net.set_train_data(x_cm, y_cm, c.m)
    d = b2nn.make_desc(momentum=c.momentum, rate=c.rate, iters=c.iters, display=c.display, batch_num=c.batch_num,
                       classify=True, act=b2nn.ACT_TANH, precision=b2nn.PREC_AUTO)
    out = net.train(d)
    ref = oracle.train(c.layers, x_cm, y_cm, c.m, batch_num=c.batch_num, iters=c.iters, display=c.display, momentum=c.momentum,
                       rate=c.rate, hidden_act=oracle.TANH, classify=True, out_act=oracle.OUT_SOFTMAX,
                       cost=oracle.COST_CROSSENTROPY, theta0=th0, precision="f64")
    assert out.gemm_tc_launches > 0, "fp32-grade default still runs on the tensor cores"
    assert np.abs(net.get_weights() - ref.theta).max() <= 2e-4 * np.abs(ref.theta).max()
    np.testing.assert_allclose(out.log_J, ref.log_J, rtol=2e-4)
    net.close()
