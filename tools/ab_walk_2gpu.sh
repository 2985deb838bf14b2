#!/bin/bash
# 2-GPU A/B of the tile walk under the peer-memory exchange (run under `gpurun --gpus 2`), then the data-parallel tests
# and the full N = 2 bench line at the shipped default.
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $1 bench.py --gpus 2 "${@:2}"; }
for rep in 1 2; do
  for gm in 16 8; do
    B2NN_TC_GROUP_M=$gm run 2960$rep --no-secondary --steps 40 --warmup 5 > gpurun_out/ab_walk_$gm.json 2> gpurun_out/ab_walk.err
    python - <<PY
import json
d = json.loads(open("gpurun_out/ab_walk_$gm.json").read().strip().splitlines()[-1])
print("group_m=$gm", d["ms_per_step"], d["e2e"]["ms_per_step"], d["replicas_identical"], {k: v["ms_per_step"] for k, v in d["roofline"]["by_kind"].items()})
PY
  done
done
timeout 500 python -m pytest tests/test_dp_gpu.py -m gpu -q > gpurun_out/r2_pytest_dp_2gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2_pytest_dp_2gpu.log
run 29611 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("gpurun_out/r2_bench_n2.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], d["replicas_identical"], d["dp_check"])
PY
