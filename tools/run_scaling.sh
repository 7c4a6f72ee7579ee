#!/bin/bash
# Multi-GPU measurements for one world size: bench.py (configs[1], one pair per GPU), all-pairs
# (configs[4], pair-sharded) and the train-sharded single pair (configs[3]).
#   gpurun --gpus N -- tools/run_scaling.sh N
# Every launch sits under its own `timeout`: a hung collective must not burn N x the box time.
N=${1:-2}
WHAT=${2:-all}
OUT=gpurun_out/scale_${N}.jsonl
mkdir -p gpurun_out; : > $OUT
if [ "$N" = "1" ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533"; fi
[ "$WHAT" = "all" ] && timeout 150 $L bench.py --gpus $N --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2> gpurun_out/scale_${N}_bench.err | grep '^{' >> $OUT
timeout 200 $L tools/bench_allpairs.py --pinned-out 2> gpurun_out/scale_${N}_allpairs.err | grep '^{' >> $OUT
timeout 120 $L tools/bench_sharded.py --size 200000 --reps 3 2> gpurun_out/scale_${N}_sharded.err | grep '^{' >> $OUT
[ "$WHAT" = "all" ] && timeout 120 $L tools/bench_sharded.py --size 200000 --dist C --reps 3 2>> gpurun_out/scale_${N}_sharded.err | grep '^{' >> $OUT
[ "$WHAT" = "all" ] && timeout 120 $L tools/bench_sharded.py --size 200000 --dist C --mode knn --reps 3 2>> gpurun_out/scale_${N}_sharded.err | grep '^{' >> $OUT
cat $OUT
