"""CPU tests of the oracle itself (no GPU): the restatement must be a correct gradient-descent loop, agree between
fp32 and fp64, partition batches like allocVarGPU, and — once generated — reproduce the reference's golden vectors."""
import os

import numpy as np
import pytest

import cases


def _run(O, c, precision="f32", theta0=None, **over):
    x_cm, y_cm = cases.oracle_inputs(c)
    kw = dict(batch_num=c.batch_num, iters=c.iters, display=c.display, momentum=c.momentum, rate=c.rate,
              hidden_act=c.hidden_act, classify=c.classify, out_act=c.out_act, cost=c.cost, precision=precision)
    kw.update(over)
    if theta0 is None:
        theta0 = O.init_weights(c.layers, c.init_kind, c.seed)
    return O.train(c.layers, x_cm, y_cm, c.m, theta0=theta0, **kw)


def test_gradient_matches_finite_differences(oracle):
    """One plain gradient step (mu = 0) must move theta by -rate * dJ/dtheta; dJ/dtheta from central differences of
    the oracle's own final-cost evaluation (cublasNN.cpp:725-745)."""
    O = oracle
    for name in ("clf_tanh_softmax", "clf_sigmoid_sigout", "reg_tanh_ragged"):
        c = cases.case_by_name(name)
        rng = np.random.default_rng(0)
        th0 = (rng.standard_normal(O.theta_size(c.layers)) * 0.3).astype(np.float32)
        step = _run(O, c, "f64", th0, iters=1, batch_num=1, momentum=0.0, rate=1.0)
        g = th0.astype(np.float64) - step.theta.astype(np.float64)
        idx = rng.choice(len(th0), size=12, replace=False)
        for i in idx:
            tp, tm = th0.copy(), th0.copy()
            tp[i] += 1e-2
            tm[i] -= 1e-2
            jp = _run(O, c, "f64", tp, iters=0, batch_num=1).final_cost
            jm = _run(O, c, "f64", tm, iters=0, batch_num=1).final_cost
            num = (jp - jm) / (float(tp[i]) - float(tm[i]))
            assert abs(num - g[i]) < 5e-5 + 2e-3 * abs(num), (name, i, num, g[i])


def test_fp32_tracks_fp64(oracle):
    for c in cases.small_cases():
        a = _run(oracle, c, "f32")
        b = _run(oracle, c, "f64")
        scale = np.abs(b.theta).max()
        assert np.abs(a.theta - b.theta).max() <= 2e-3 * scale, c.name
        np.testing.assert_allclose(a.log_J, b.log_J, rtol=2e-3)
        assert abs(a.final_cost - b.final_cost) <= 2e-3 * abs(b.final_cost) + 1e-6


def test_cost_decreases(oracle):
    for c in cases.small_cases():
        r = _run(oracle, c, "f64")
        assert len(r.log_J) == c.iters // c.display
        assert r.log_J[-1] < r.log_J[0], (c.name, r.log_J)
        assert list(r.log_iter) == [i for i in range(c.iters) if (i + 1) % c.display == 0]


def test_batch_partition_last_batch_absorbs_remainder(oracle):
    """batchNum is a COUNT; rows are contiguous; the last batch takes m - (batchNum-1)*floor(m/batchNum) rows
    (cublasNN.cpp:401, 421-432).  With mu = 0 and a tiny rate, one epoch over b batches must equal the sum of the
    per-batch gradient steps taken at (almost) the same theta: check instead the exact property that permuting rows
    INSIDE a batch leaves the result unchanged while moving a row ACROSS the boundary changes it."""
    O = oracle
    c = cases.case_by_name("reg_tanh_ragged")          # m = 205, 7 batches: 6 x 29 + 31
    base = _run(O, c, "f64", iters=1)
    x2, y2 = c.x.copy(), c.y.copy()
    x2[[0, 5]] = x2[[5, 0]]
    y2[[0, 5]] = y2[[5, 0]]                            # rows 0 and 5 are both in batch 0
    c2 = cases.Case(**{**c.__dict__, "x": x2, "y": y2})
    same = _run(O, c2, "f64", iters=1)
    np.testing.assert_allclose(same.theta, base.theta, rtol=0, atol=1e-6)
    x3, y3 = c.x.copy(), c.y.copy()
    x3[[28, 29]] = x3[[29, 28]]
    y3[[28, 29]] = y3[[29, 28]]                        # row 28 is the last of batch 0, row 29 the first of batch 1
    c3 = cases.Case(**{**c.__dict__, "x": x3, "y": y3})
    moved = _run(O, c3, "f64", iters=1)
    assert np.abs(moved.theta - base.theta).max() > 1e-6
    # rows 174..204 (31 rows) form the last batch: swapping 174 and 204 is again invisible
    x4, y4 = c.x.copy(), c.y.copy()
    x4[[174, 204]] = x4[[204, 174]]
    y4[[174, 204]] = y4[[204, 174]]
    c4 = cases.Case(**{**c.__dict__, "x": x4, "y": y4})
    last = _run(O, c4, "f64", iters=1)
    np.testing.assert_allclose(last.theta, base.theta, rtol=0, atol=1e-6)


def test_init_weights_ranges_and_determinism(oracle):
    O = oracle
    layers = [784, 50, 10]
    a = O.init_weights(layers, 2, 42)
    b = O.init_weights(layers, 2, 42)
    c = O.init_weights(layers, 2, 43)
    assert np.array_equal(a, b) and not np.array_equal(a, c)
    for (pos, r, cols), kind in zip(O.theta_offsets(layers), (2, 2)):
        eps = np.sqrt(6.0) / np.sqrt(r + cols)
        blk = a[pos:pos + r * cols]
        assert blk.min() >= -eps * (1 + 1e-6) and blk.max() <= eps * (1 + 1e-6)
        assert abs(blk.mean()) < 0.05 * eps
    d = O.init_weights([10, 40, 160, 1], 1, 42)
    pos, r, cols = O.theta_offsets([10, 40, 160, 1])[1]
    assert np.abs(d[pos:pos + r * cols]).max() <= 1 / np.sqrt(41) * (1 + 1e-6)


def test_predict_matches_training_forward(oracle):
    O = oracle
    c = cases.case_by_name("clf_tanh_softmax")
    r = _run(O, c, "f32")
    x_cm, _ = cases.oracle_inputs(c)
    cls, probs = O.predict(c.layers, r.theta, x_cm, c.m, hidden_act=c.hidden_act, classify=True, out_act=c.out_act)
    probs = probs.reshape(c.layers[-1], c.m)
    np.testing.assert_allclose(probs.sum(axis=0), 1.0, rtol=1e-5)
    assert np.array_equal(cls, probs.argmax(axis=0).astype(np.float32))
    errs = float((cls != c.y[:, 0]).sum())
    assert errs == r.error_count


@pytest.mark.parametrize("name", [c.name for c in cases.small_cases()])
def test_oracle_reproduces_reference_golden(oracle, name):
    """The pin: outputs of the reference's own cuBLAS build on a B200 (tests/golden/make_golden.py)."""
    path = cases.golden_path(name)
    if not os.path.exists(path):
        pytest.skip("golden vectors not generated yet (needs one GPU run of tests/golden/make_golden.py)")
    g = np.load(path)
    c = cases.case_by_name(name)
    th0 = oracle.init_weights(c.layers, c.init_kind, c.seed)
    assert np.array_equal(th0, g["theta0"]), "cuRAND host generator + rescale must be bit-identical to setLayers"
    r = _run(oracle, c, "f32", g["theta0"])
    scale = np.abs(g["theta"]).max()
    assert np.abs(r.theta - g["theta"]).max() <= 1e-4 * scale
    if c.compare_log:
        np.testing.assert_allclose(r.log_J, g["log_J"], rtol=2e-4)
        assert list(r.log_iter) == list(g["log_iter"])
    assert abs(r.final_cost - float(g["final_cost"])) <= 2e-4 * abs(float(g["final_cost"])) + 1e-7
    if c.classify:
        assert abs(r.error_count - float(g["error_count"])) <= max(1.0, 0.002 * c.m)
