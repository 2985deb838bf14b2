#!/bin/bash
# One multi-GPU session (run under `gpurun --gpus 8`): the data-parallel parity tests at 2 / 4 / 8 ranks, then bench.py
# at N = 8 and N = 4 — the evidence behind the scaling rows of BASELINE.md.  Logs land in gpurun_out/ (copied to profiles/).
nvidia-smi -L | wc -l
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2_smoke.log
timeout 600 python -m pytest tests/test_dp_gpu.py -m gpu -q -k "c4_mini or custom" > gpurun_out/r2_pytest_dp_8gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r2_pytest_dp_8gpu.log
tail -3 gpurun_out/r2_pytest_dp_8gpu.log
for n in 8 4; do
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2952$n bench.py --gpus $n > gpurun_out/r2_bench_n$n.json 2> gpurun_out/r2_bench_n$n.err
    echo "n=$n rc=$?"
    python - <<PY
import json
p = json.loads(open("gpurun_out/r2_bench_n$n.json").read().strip().splitlines()[-1])
print(p["value"], p["ms_per_step"], p["replicas_identical"], p["e2e"]["ms_per_step"], {k: x["ms_per_step"] for k, x in p["roofline"]["by_kind"].items()})
print(p["dp_check"])
c = p["secondary"]["c4"]
print({k: c.get(k) for k in ("ms_per_step", "samples_per_s", "tflops", "replicas_identical", "gradient_exchange")}, c.get("roofline", {}).get("frac"))
PY
done
