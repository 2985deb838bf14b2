#!/usr/bin/env python
"""Tile-walk sweep for the big contractions: sustained ms/step of config C3 (784-4096-4096-10, batch
8192, plain tf32) for each rasterisation group size (B2NN_TC_GROUP_M) with and without the serpentine column order
(B2NN_TC_SERP).  Both knobs are read at launch time, so one process covers the whole sweep.

    python tools/sweep_tile_walk.py [--steps 1000] [--short] [--ab] [--precision tf32|tf32x3|bf16]      # --short: 3 steps per setting (for an ncu metrics pass)
"""
from __future__ import annotations

import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from measure_configs import load_b2nn, synth_classify  # noqa: E402


def main():
    steps = 1000
    if "--steps" in sys.argv:
        steps = int(sys.argv[sys.argv.index("--steps") + 1])
    short = "--short" in sys.argv
    prec_name = sys.argv[sys.argv.index("--precision") + 1] if "--precision" in sys.argv else "tf32"
    b2nn = load_b2nn()
    layers = [784, 4096, 4096, 10]
    m = 8192
    x, y, _ = synth_classify(layers, m)
    settings = [(16, 0), (4, 0), (8, 0), (8, 1), (16, 1), (32, 0), (32, 1), (2, 1), (4, 1), (16, 0)]
    if short:
        settings = [(16, 0), (8, 0), (8, 1), (4, 1), (32, 1)]
    if "--ab" in sys.argv:
        settings = [(16, 0), (8, 1), (16, 0), (8, 1)]
    prec = {"tf32": b2nn.PREC_TF32, "tf32x3": b2nn.PREC_TF32X3, "bf16": b2nn.PREC_BF16}[prec_name]
    for gm, serp in settings:
        os.environ["B2NN_TC_GROUP_M"] = str(gm)
        os.environ["B2NN_TC_SERP"] = str(serp)
        net = b2nn.Net(0)
        net.set_layers(layers, b2nn.INIT_2, 42)
        net.set_train_data(x.reshape(-1), y.reshape(-1), m)
        kw = dict(momentum=0.9, rate=0.05, display=1 << 30, batch_num=1, classify=True, act=b2nn.ACT_SIGMOID,
                  out=b2nn.OUT_SOFTMAX, cost=b2nn.COST_CROSSENTROPY, precision=prec, skip_final_cost=True)
        if short:
            out = net.train(b2nn.make_desc(iters=3, **kw))
        else:
            net.train(b2nn.make_desc(iters=100, **kw))
            out = net.train(b2nn.make_desc(iters=steps, **kw))
        print(json.dumps({"precision": prec_name, "group_m": gm, "serp": serp, "steps": out.steps,
                          "ms_per_step": round(out.loop_seconds * 1e3 / out.steps, 4)}), flush=True)
        net.close()


if __name__ == "__main__":
    main()
