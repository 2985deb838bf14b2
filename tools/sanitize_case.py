#!/usr/bin/env python
"""Train one parity case through the C ABI — the workload compute-sanitizer is pointed at (tools/sanitize.sh).

    python tools/sanitize_case.py <case> <fp32|tf32|tf32x3> [--dp]      (--dp: under torchrun, 2+ ranks, peer-memory exchange)
"""
import importlib.util
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    import cases
    b2nn = load("b2nn_binding", os.path.join(ROOT, "cublas-ml_b200", "b2nn.py"))
    name, prec = sys.argv[1], {"fp32": 0, "tf32": 1, "tf32x3": 3}[sys.argv[2]]
    dp_mode = "--dp" in sys.argv
    c = cases.case_by_name(name)
    x_cm, y_cm = cases.oracle_inputs(c)
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    net = b2nn.Net(int(os.environ.get("LOCAL_RANK", "0")))
    net.set_layers(c.layers, c.init_kind, c.seed)
    m = c.m
    if dp_mode and world > 1:
        import torch
        import torch.distributed as dist
        dp = load("b2nn_dp", os.path.join(ROOT, "cublas-ml_b200", "dp.py"))
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("gloo")
        net.comm_init(rank, world, dp.exchange_unique_id(b2nn.Net.comm_unique_id, rank, world))
        dp.enable_peer_exchange(net, rank, world)
        x_cm, y_cm, m = dp.shard_training_set(x_cm, y_cm, c.m, c.layers[0] + 1, c.layers[-1], c.batch_num, rank, world)
    net.set_train_data(x_cm, y_cm, m)
    d = b2nn.make_desc(momentum=c.momentum, rate=c.rate, iters=min(c.iters, 3), display=1, batch_num=c.batch_num,
                       classify=c.classify, act=b2nn.ACT_TANH if c.hidden_act == cases.TANH else b2nn.ACT_SIGMOID,
                       out=b2nn.OUT_SOFTMAX if c.classify else b2nn.OUT_LINEAR,
                       cost=b2nn.COST_CROSSENTROPY if c.classify else b2nn.COST_SQUARED, precision=prec)
    out = net.train(d)
    got, _ = net.predict(d, cases.oracle_inputs(c)[0], c.m)
    print(f"rank {rank}: case {name} prec {sys.argv[2]} J={out.log_J.tolist()} final={out.final_cost:.6f} "
          f"launches={out.kernel_launches} tcgen05={out.gemm_tc_launches} pred_sum={float(np.sum(got)):.0f}", flush=True)
    net.close()


if __name__ == "__main__":
    main()
