"""Golden vectors: outputs of the UNMODIFIED reference (oracle/_ref/libcublasml_ref.so, built from /root/reference
by oracle/Makefile) on the seeded cases of tests/cases.py.  Needs a GPU:

    gpurun -- 'python tests/golden/make_golden.py'      # writes gpurun_out/golden/ref_<case>.npz
    cp gpurun_out/golden/ref_*.npz tests/golden/         # commit them

Each file holds theta0 (the reference's cuRAND draw for the case's seed), theta (after training), the printed
(iteration, cost) pairs, the final cost and the error count — the anchors the CPU oracle and the CUDA engine are
checked against (tests/test_oracle_golden.py, tests/test_train_gpu.py).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cases  # noqa: E402
from oracle import oracle as O  # noqa: E402


def main():
    out_dir = os.path.join(ROOT, "gpurun_out", "golden")
    os.makedirs(out_dir, exist_ok=True)
    for c in cases.small_cases():
        r = O.ref_train(c.layers, c.x, c.y, batch_num=c.batch_num, iters=c.iters, display=c.display,
                        momentum=c.momentum, rate=c.rate, hidden_act=c.hidden_act, classify=c.classify,
                        out_act=c.out_act, cost=c.cost, init_kind=c.init_kind, seed=c.seed)
        np.savez_compressed(os.path.join(out_dir, f"ref_{c.name}.npz"), theta0=r.theta0, theta=r.theta,
                            log_iter=r.log_iter, log_J=r.log_J, final_cost=np.float64(r.final_cost),
                            error_count=np.float64(r.error_count))
        print(c.name, "J", r.log_J[:3], "...", r.log_J[-1:] if len(r.log_J) else None, "final", r.final_cost,
              "errors", r.error_count, flush=True)


if __name__ == "__main__":
    main()
