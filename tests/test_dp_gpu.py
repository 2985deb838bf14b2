"""Data-parallel engine parity on real GPUs (needs >= 2; skipped otherwise): a 2-rank run, each rank holding its row
slice of every batch with gradients all-reduced by NCCL inside the engine, must walk the same trajectory as one GPU
training on the whole batches.  Summation order differs, so this is a tolerance check."""
import importlib.util
import os
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, name, prec_name, p2p, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    from conftest import load_b2nn
    import cases
    spec = importlib.util.spec_from_file_location("b2nn_dp", os.path.join(ROOT, "cublas-ml_b200", "dp.py"))
    dp = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(dp)
    b2nn = load_b2nn()
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(rank)
        c = cases.case_by_name(name)
        x_cm, y_cm = cases.oracle_inputs(c)
        n1, K = c.layers[0] + 1, c.layers[-1]
        net = b2nn.Net(rank)
        net.set_layers(c.layers, c.init_kind, c.seed)
        net.comm_init(rank, world, dp.exchange_unique_id(b2nn.Net.comm_unique_id, rank, world))
        if p2p:
            assert dp.enable_peer_exchange(net, rank, world)
        xl, yl, ml = dp.shard_training_set(x_cm, y_cm, c.m, n1, K, c.batch_num, rank, world)
        net.set_train_data(xl, yl, ml)
        d = b2nn.make_desc(momentum=c.momentum, rate=c.rate, iters=c.iters, display=c.display, batch_num=c.batch_num,
                           classify=c.classify, act=b2nn.ACT_TANH if c.hidden_act == cases.TANH else b2nn.ACT_SIGMOID,
                           out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY,
                           precision=b2nn.PREC_TF32 if prec_name == "tf32" else b2nn.PREC_FP32)
        out = net.train(d)
        th = net.get_weights()
        net.comm_barrier()
        q.put((rank, th, out.log_J, out.final_cost, out.error_count, out.kernel_launches))
        net.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("name,prec,p2p", [("clf_tanh_softmax", "fp32", False), ("wide_mini", "fp32", False),
                                           ("wide_mini", "tf32", False), ("wide_mini", "tf32", True),
                                           ("wide1024_skinny", "tf32", True), ("wide1024_skinny", "tf32", False)])
def test_two_gpu_run_matches_single_gpu(b2nn, name, prec, p2p):
    """p2p = True: the first layer's gradient (512 rows = 2 ranks x 256-column tiles) takes the fused peer-memory path
    (reduce-scatter in the GEMM epilogue, update + all-gather on the owner); the others stay on NCCL."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    import cases
    c = cases.case_by_name(name)
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + (os.getpid() % 300)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, name, prec, p2p, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-GPU reference run of the same global problem
    net = b2nn.Net(0)
    net.set_layers(c.layers, c.init_kind, c.seed)
    x_cm, y_cm = cases.oracle_inputs(c)
    net.set_train_data(x_cm, y_cm, c.m)
    d = b2nn.make_desc(momentum=c.momentum, rate=c.rate, iters=c.iters, display=c.display, batch_num=c.batch_num,
                       classify=c.classify, act=b2nn.ACT_TANH if c.hidden_act == cases.TANH else b2nn.ACT_SIGMOID,
                       out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY,
                       precision=b2nn.PREC_TF32 if prec == "tf32" else b2nn.PREC_FP32)
    one = net.train(d)
    th1 = net.get_weights()
    net.close()
    tol = 2e-4 if prec == "fp32" else 2e-2
    assert np.array_equal(res[0][1], res[1][1]), "replicas diverged: every rank must hold identical weights"
    assert np.abs(res[0][1] - th1).max() <= tol * np.abs(th1).max()
    for r in res:
        np.testing.assert_allclose(r[2], one.log_J, rtol=tol)          # printed cost = GLOBAL batch-0 cost on every rank
        assert abs(r[3] - one.final_cost) <= tol * abs(one.final_cost) + 1e-7
        assert abs(r[4] - one.error_count) <= max(1.0, 0.005 * c.m)
