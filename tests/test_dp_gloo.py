"""Data-parallel host logic on CPU: world_size 2 over gloo.  Checks the sharding helper, the unique-id exchange, and
the identity the engine's gradient all-reduce relies on: summing the per-rank updates (each scaled by
rate / GLOBAL batch) reproduces the single-process step (cublasNN.cpp:601 alpha, :680-684 Delta = delta^T a)."""
import importlib.util
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _load_dp():
    spec = importlib.util.spec_from_file_location("b2nn_dp", os.path.join(ROOT, "cublas-ml_b200", "dp.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_batch_table_and_slices():
    dp = _load_dp()
    assert dp.batch_table(205, 7) == [(0, 29), (29, 58), (58, 87), (87, 116), (116, 145), (145, 174), (174, 205)]
    for lo, hi, world in ((0, 29, 2), (174, 205, 4), (0, 8192, 8), (5, 6, 3)):
        parts = [dp.rank_slice(lo, hi, r, world) for r in range(world)]
        assert parts[0][0] == lo and parts[-1][1] == hi
        assert all(parts[i][1] == parts[i + 1][0] for i in range(world - 1))
        sizes = [e - s for s, e in parts]
        assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle import oracle as O
    import cases
    dp = _load_dp()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        uid = dp.exchange_unique_id(lambda: bytes(range(128)), rank, world)
        assert uid == bytes(range(128))

        c = cases.case_by_name("clf_tanh_softmax")          # m = 300, 4 batches of 75 -> use 2 batches of 150
        nb = 2
        x_cm, y_cm = cases.oracle_inputs(c)
        n1, K, m = c.layers[0] + 1, c.layers[-1], c.m
        th0 = O.init_weights(c.layers, c.init_kind, c.seed)
        xl, yl, ml = dp.shard_training_set(x_cm, y_cm, m, n1, K, nb, rank, world)
        assert ml == m // world
        kw = dict(display=1, momentum=0.0, hidden_act=c.hidden_act, classify=True, out_act=c.out_act, cost=c.cost,
                  precision="f64")
        # one step on batch 0 only (iters = 1, batch_num = 1 over the rank's slice of batch 0)
        lo, hi = dp.batch_table(m, nb)[0]
        s, e = dp.rank_slice(lo, hi, rank, world)
        X = x_cm.reshape(n1, m)[:, s:e]
        Y = y_cm.reshape(K, m)[:, s:e]
        rows = e - s
        rate = 0.4
        # the oracle divides by the LOCAL rows; rescale so the step uses rate / GLOBAL rows like the engine does
        local = O.train(c.layers, np.ascontiguousarray(X).reshape(-1), np.ascontiguousarray(Y).reshape(-1), rows,
                        batch_num=1, iters=1, rate=rate * rows / (hi - lo), theta0=th0, **kw)
        delta = torch.tensor(local.theta.astype(np.float64) - th0.astype(np.float64))
        dist.all_reduce(delta)                               # what ncclAllReduce(G) + the update amount to
        Xg = x_cm.reshape(n1, m)[:, lo:hi]
        Yg = y_cm.reshape(K, m)[:, lo:hi]
        full = O.train(c.layers, np.ascontiguousarray(Xg).reshape(-1), np.ascontiguousarray(Yg).reshape(-1), hi - lo,
                       batch_num=1, iters=1, rate=rate, theta0=th0, **kw)
        err = float(np.abs(th0 + delta.numpy() - full.theta).max())
        q.put((rank, err, float(np.abs(full.theta - th0).max())))
    finally:
        dist.destroy_process_group()


def test_dp_identity_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29400 + (os.getpid() % 500)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    results = [q.get(timeout=5) for _ in range(2)]
    for rank, err, step in results:
        assert step > 1e-4
        assert err <= 2e-6 * max(1.0, step) + 3e-7, (rank, err, step)     # theta is returned as float32


def test_shards_partition_every_batch():
    """All ranks' local sets together hold every row of every batch exactly once, in batch order, and each local set
    is cut by the engine's own batch table (floor(m_local / batch_num), last absorbs the rest) into those slices."""
    dp = _load_dp()
    rng = np.random.default_rng(0)
    for m, nb, world in ((10, 3, 2), (205, 7, 4), (64, 1, 8), (4200 * 3, 3, 8)):
        n1, K = 3, 2
        x = rng.standard_normal(n1 * m).astype(np.float32)
        y = rng.standard_normal(K * m).astype(np.float32)
        X = x.reshape(n1, m)
        seen = []
        for r in range(world):
            xl, yl, ml = dp.shard_training_set(x, y, m, n1, K, nb, r, world)
            assert xl.size == n1 * ml and yl.size == K * ml
            local_tab = dp.batch_table(ml, nb)
            for (lo, hi), (llo, lhi) in zip(dp.batch_table(m, nb), local_tab):
                s, e = dp.rank_slice(lo, hi, r, world)
                assert lhi - llo == e - s
                assert np.array_equal(xl.reshape(n1, ml)[:, llo:lhi], X[:, s:e])
                seen.extend(range(s, e))
        assert sorted(seen) == list(range(m))
