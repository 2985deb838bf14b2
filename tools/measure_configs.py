#!/usr/bin/env python
"""Measure the five BASELINE.json configs on one B200: this engine (tf32 and fp32 modes) next to the unmodified
reference (oracle/_ref, cuBLAS fp32) — the numbers behind BASELINE.md §5 / DESIGN.md §5.  Writes one JSON object per
line to stdout.  Synthetic inputs of each config's shape (C2 uses the committed bikeshare fixture).

    python tools/measure_configs.py [c1 c2 c3 c4 c5] [--no-ref]
"""
from __future__ import annotations

import ctypes
import importlib.util
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def load_b2nn():
    spec = importlib.util.spec_from_file_location("b2nn_binding", os.path.join(ROOT, "cublas-ml_b200", "b2nn.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["b2nn_binding"] = mod
    spec.loader.exec_module(mod)
    return mod


def step_flops(layers, B):
    fwd = sum(2.0 * B * (layers[j] + 1) * layers[j + 1] for j in range(len(layers) - 1))
    bwd = sum(2.0 * B * layers[j] * layers[j + 1] for j in range(1, len(layers) - 1))
    return 2 * fwd + bwd


def synth_classify(layers, m, seed=1):
    import torch
    n, K = layers[0], layers[-1]
    g = torch.Generator().manual_seed(seed)
    x = torch.empty(n + 1, m)
    x[0] = 1.0
    x[1:].normal_(generator=g)
    lab = (torch.randn(K, n, generator=torch.Generator().manual_seed(2)) @ x[1:]).argmax(0)
    y = torch.zeros(K, m)
    y[lab, torch.arange(m)] = 1.0
    return x.numpy(), y.numpy(), lab.numpy()


def ours_train(b2nn, layers, x, y, m, nb, iters, classify, act, prec, rate):
    net = b2nn.Net(0)
    net.set_layers(layers, b2nn.INIT_2 if classify else b2nn.INIT_1, 42)
    net.set_train_data(x.reshape(-1), y.reshape(-1), m)
    d = b2nn.make_desc(momentum=0.9, rate=rate, iters=iters, display=1 << 30, batch_num=nb, classify=classify, act=act,
                       out=b2nn.OUT_SOFTMAX if classify else b2nn.OUT_LINEAR,
                       cost=b2nn.COST_CROSSENTROPY if classify else b2nn.COST_SQUARED, precision=prec,
                       skip_final_cost=True)
    warm = b2nn.make_desc(momentum=0.9, rate=rate, iters=max(1, 3 // nb + 1), display=1 << 30, batch_num=nb, classify=classify,
                          act=act, out=d.out_kind, cost=d.cost_kind, precision=prec, skip_final_cost=True)
    net.train(warm)
    out = net.train(d)
    if os.environ.get("B2NN_MEASURE_PROFILE"):
        net.profile_enable(True)
        net.train(d)
        prof = net.profile_read()
        net.profile_enable(False)
        steps = out.steps
        print(json.dumps({"profile_us_per_step": {k: round(v["ms"] * 1e3 / steps, 2) for k, v in prof.items()},
                          "launches_per_step": {k: v["launches"] / steps for k, v in prof.items()},
                          "tc_per_step": {k: v["tc_launches"] / steps for k, v in prof.items()}}), flush=True)
    net.close()
    return out


def ref_train_time(O, layers, x_rows, y_rows, nb, it_a, it_b, classify, hidden, rate, init_kind):
    kw = dict(batch_num=nb, display=1 << 30, momentum=0.9, rate=rate, hidden_act=hidden, classify=classify,
              out_act=O.OUT_SOFTMAX, cost=O.COST_CROSSENTROPY, init_kind=init_kind, seed=42)
    O.ref_train(layers, x_rows, y_rows, iters=it_a, **kw)
    ta = O.ref_train(layers, x_rows, y_rows, iters=it_a, **kw).seconds
    tb = O.ref_train(layers, x_rows, y_rows, iters=it_b, **kw).seconds
    return max(tb - ta, 1e-9) / ((it_b - it_a) * nb)


def emit(**kw):
    print(json.dumps(kw), flush=True)


def main():
    import torch
    from oracle import oracle as O
    import cases
    b2nn = load_b2nn()
    which = [a for a in sys.argv[1:] if not a.startswith("-")] or ["c1", "c2", "c3", "c4", "c5"]
    do_ref = "--no-ref" not in sys.argv and O.ref_available()

    if "c1" in which:
        layers, m, nb = [784, 50, 10], 42000, 10
        lab = cases.mnist_labels()
        xr = cases.zscore(cases.synth_digits(lab, 784, seed=1234))
        x, y = O.with_bias_colmajor(xr), O.one_hot_colmajor(lab, 10)
        for prec, name in ((b2nn.PREC_TF32, "tf32"), (b2nn.PREC_FP32, "fp32")):
            out = ours_train(b2nn, layers, x, y, m, nb, 30, True, b2nn.ACT_TANH, prec, 0.075)
            us = out.loop_seconds / out.steps * 1e6
            emit(config="C1", impl="ours", mode=name, us_per_step=round(us, 2), samples_per_s=round(4200 / us * 1e6),
                 launches_per_step=out.kernel_launches / out.steps, hbm_bound_us=2.16)
        if do_ref:
            t = ref_train_time(O, layers, xr, lab.astype(np.float32).reshape(-1, 1), nb, 10, 110, True, O.TANH, 0.075, 2)
            emit(config="C1", impl="reference", mode="fp32", us_per_step=round(t * 1e6, 2), samples_per_s=round(4200 / t))

    if "c2" in which:
        b = cases.bikeshare()
        layers, m = [10, 40, 160, 1], 4337
        xr = cases.zscore(b["x"])
        x, y = O.with_bias_colmajor(xr), O.colmajor(b["y"])
        for prec, name in ((b2nn.PREC_TF32, "tf32"), (b2nn.PREC_FP32, "fp32")):
            out = ours_train(b2nn, layers, x, y, m, 1, 3000, False, b2nn.ACT_SIGMOID, prec, 0.003)
            us = out.loop_seconds / out.steps * 1e6
            emit(config="C2", impl="ours", mode=name, us_per_step=round(us, 2), samples_per_s=round(m / us * 1e6),
                 launches_per_step=out.kernel_launches / out.steps)
        if do_ref:
            t = ref_train_time(O, layers, xr, b["y"], 1, 500, 3500, False, O.SIGMOID, 0.003, 1)
            emit(config="C2", impl="reference", mode="fp32", us_per_step=round(t * 1e6, 2), samples_per_s=round(m / t))

    if "c3" in which:
        layers, B, nb = [784, 4096, 4096, 10], 8192, 8
        x, y, lab = synth_classify(layers, B * nb)
        for prec, name in ((b2nn.PREC_TF32, "tf32"), (b2nn.PREC_FP32, "fp32")):
            out = ours_train(b2nn, layers, x, y, B * nb, nb, 2 if name == "tf32" else 1, True, b2nn.ACT_TANH, prec, 0.01)
            ms = out.loop_seconds / out.steps * 1e3
            emit(config="C3", impl="ours", mode=name, ms_per_step=round(ms, 4), samples_per_s=round(B / ms * 1e3),
                 tflops=round(step_flops(layers, B) / ms / 1e9, 1))
        if do_ref:
            t = ref_train_time(O, layers, np.ascontiguousarray(x[1:].T), lab.astype(np.float32).reshape(-1, 1), nb, 1, 3,
                               True, O.TANH, 0.01, 2)
            emit(config="C3", impl="reference", mode="fp32", ms_per_step=round(t * 1e3, 4), samples_per_s=round(B / t),
                 tflops=round(step_flops(layers, B) / t / 1e12, 1))

    if "c4" in which:
        layers = [784] + [8192] * 8 + [10]
        for B in (8192, 16384):
            nb = 2
            x, y, lab = synth_classify(layers, B * nb)
            out = ours_train(b2nn, layers, x, y, B * nb, nb, 2, True, b2nn.ACT_TANH, b2nn.PREC_TF32, 0.01)
            ms = out.loop_seconds / out.steps * 1e3
            emit(config="C4", impl="ours", mode="tf32", batch=B, ms_per_step=round(ms, 3), samples_per_s=round(B / ms * 1e3),
                 tflops=round(step_flops(layers, B) / ms / 1e9, 1))
            if do_ref and B == 8192:
                t = ref_train_time(O, layers, np.ascontiguousarray(x[1:].T), lab.astype(np.float32).reshape(-1, 1), nb, 1,
                                   2, True, O.TANH, 0.01, 2)
                emit(config="C4", impl="reference", mode="fp32", batch=B, ms_per_step=round(t * 1e3, 3),
                     samples_per_s=round(B / t), tflops=round(step_flops(layers, B) / t / 1e12, 1))

    if "c5" in which:
        layers = [784, 50, 10]
        net = b2nn.Net(0)
        net.set_layers(layers, b2nn.INIT_2, 42)
        stream = torch.cuda.ExternalStream(net.stream())
        for prec, name in ((b2nn.PREC_TF32, "tf32"), (b2nn.PREC_FP32, "fp32")):
            d = b2nn.make_desc(classify=True, act=b2nn.ACT_TANH, out=b2nn.OUT_SOFTMAX, precision=prec)
            for p in range(0, 21, 2):
                Bn = 1 << p
                ld = (Bn + 31) // 32 * 32
                xd = torch.randn(785, ld, device="cuda")
                xd[784] = 1.0
                outd = torch.empty(ld, device="cuda")
                for _ in range(3):
                    net.predict_device(d, xd.data_ptr(), Bn, ld, outd.data_ptr())
                reps = 20 if Bn <= 65536 else 5
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(stream)
                for _ in range(reps):
                    net.predict_device(d, xd.data_ptr(), Bn, ld, outd.data_ptr())
                e1.record(stream)
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / reps * 1e3
                emit(config="C5", impl="ours", mode=name, batch=Bn, us=round(us, 2), samples_per_s=round(Bn / us * 1e6),
                     hbm_gbs=round(Bn * 3144 / us / 1e3, 1))
                if os.environ.get("B2NN_MEASURE_PROFILE") and Bn >= 65536:
                    net.profile_enable(True)
                    for _ in range(3):
                        net.predict_device(d, xd.data_ptr(), Bn, ld, outd.data_ptr())
                    prof = net.profile_read()
                    net.profile_enable(False)
                    print(json.dumps({"c5_profile_us": {k: round(v["ms"] * 1e3 / 3, 1) for k, v in prof.items() if v["launches"]}}), flush=True)
                # end to end through b2nn_predict (host rows in, class ids out)
                if name == "tf32" and Bn in (1, 1024, 65536, 1 << 20):
                    xh = np.random.default_rng(0).standard_normal((785, Bn)).astype(np.float32)
                    xh[0] = 1.0
                    t0 = time.perf_counter()
                    net.predict(d, xh.reshape(-1), Bn)
                    t1 = time.perf_counter()
                    emit(config="C5", impl="ours", mode="tf32 end-to-end b2nn_predict", batch=Bn, us=round((t1 - t0) * 1e6, 1),
                         samples_per_s=round(Bn / (t1 - t0)))
                del xd, outd
        net.close()


if __name__ == "__main__":
    main()
