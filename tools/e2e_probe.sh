for k in 20 100; do
  for inl in 0 1; do
    if [ $inl = 1 ]; then export B2NN_LOSS_D2H_INLINE=1; else unset B2NN_LOSS_D2H_INLINE; fi
    python bench.py --no-secondary --steps $k --warmup 5 > gpurun_out/e2e_probe_${k}_${inl}.json 2>gpurun_out/e2e_probe.err
    python - <<PY
import json
d=json.loads(open("gpurun_out/e2e_probe_${k}_${inl}.json").read().strip().splitlines()[-1])
print("K=$k inline=$inl", d["ms_per_step"], d["e2e"]["ms_per_step"], d["e2e"]["device_loop_ms_per_step"], d["clocks"])
PY
  done
done
unset B2NN_LOSS_D2H_INLINE
python tools/sweep_tile_walk.py --ab --steps 300 --precision tf32x3
python tools/sweep_tile_walk.py --ab --steps 500 --precision bf16
