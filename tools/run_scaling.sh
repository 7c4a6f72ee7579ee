#!/bin/bash
# Multi-GPU measurements for one world size: bench.py (configs[4] pair-sharded headline + the configs[3] train-sharded
# extra) and the bit-identity check of the train-sharded pair against the unsharded call.
#   gpurun --gpus N -- tools/run_scaling.sh N [steps]
# Every launch sits under its own `timeout`: a hung collective must not burn N x the box time.
N=${1:-2}
STEPS=${2:-3}
OUT=gpurun_out/scale_${N}.jsonl
mkdir -p gpurun_out; : > $OUT
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
timeout 600 $L bench.py --gpus $N --steps $STEPS --warmup 3 --no-cpu-baseline 2> gpurun_out/scale_${N}_bench.err | grep '^{' >> $OUT
timeout 300 $L tools/bench_sharded.py --size 200000 --check 2> gpurun_out/scale_${N}_sharded.err | grep '^{' >> $OUT
tail -c 3000 gpurun_out/scale_${N}_bench.err gpurun_out/scale_${N}_sharded.err
cat $OUT
