#!/bin/bash
# compute-sanitizer over the kernel families, ONE TOOL PER CALL (B200_PROFILING.md: several sanitizer tools in one gpurun
# call have wedged the GPU on this driver):   tools/sanitize.sh memcheck|racecheck|synccheck [outdir]
# The summary lands in $OUT/r2_sanitizer_<tool>.txt and is committed under profiles/.
#   memcheck : out-of-bounds / misaligned accesses;  racecheck: shared-memory hazards (mbarrier / TMA / tcgen05 pipelines,
#   the PDL rule "wait, then launch_dependents");  synccheck: divergent barriers.
#   wide_mini  tf32 / tf32x3 : tcgen05 + TMA + TMEM path, skinny / output kernels
#   mnist_mini tf32x3        : one-kernel step (mlp3.cu); with B2NN_NO_MLP3=1 the tcgen05 swapped plan + tail kernels
#   reg_sigmoid_bikeshare512 : micro kernel (fp32)
#   2 ranks, wide_mini tf32  : peer-memory exchange (flags, slot buffers, scatter epilogue)   [memcheck, needs 2 GPUs]
set -u
TOOL=${1:?tool}
OUT=${2:-gpurun_out}
SUM=$OUT/r2_sanitizer_${TOOL}.txt
S="compute-sanitizer --error-exitcode 1 --print-limit 20 --tool $TOOL"
run() {  # name cmd...
    local name=$1; shift
    echo "=== $name [$TOOL]" >> $SUM
    B2NN_GRAPH=0 timeout 600 $S "$@" > $OUT/r2_san_${name}_${TOOL}.log 2>&1
    echo "exit=$? $(grep -E 'ERROR SUMMARY|RACECHECK SUMMARY|hazard' $OUT/r2_san_${name}_${TOOL}.log | tail -3 | tr '\n' ' ')" >> $SUM
    grep -E "^rank" $OUT/r2_san_${name}_${TOOL}.log | tail -2 >> $SUM
}
echo "# compute-sanitizer --tool $TOOL, $(nvidia-smi --query-gpu=name --format=csv,noheader | head -1), $(date -u +%FT%TZ)" > $SUM
run wide_tf32 python tools/sanitize_case.py wide_mini tf32
run mnist_mlp3 python tools/sanitize_case.py mnist_mini tf32x3
run wide_x3 python tools/sanitize_case.py wide_mini tf32x3
export B2NN_NO_MLP3=1
run mnist_tail python tools/sanitize_case.py mnist_mini tf32
unset B2NN_NO_MLP3
run micro python tools/sanitize_case.py reg_sigmoid_bikeshare512 fp32
if [ "$TOOL" = memcheck ] && [ "$(nvidia-smi -L | wc -l)" -ge 2 ]; then
    run dp2_wide --target-processes all python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/sanitize_case.py wide_mini tf32 --dp
fi
cat $SUM
