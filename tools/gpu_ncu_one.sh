#!/bin/bash
# one `ncu --set full` capture of a single kernel in steady state (16,384 chains, 44 warm-up iterations).
# usage: tools/gpu_ncu_one.sh <tag> <kernel regex> [launches to skip = 44] [extra bench args]
TAG=$1; RE=$2; SKIP=${3:-44}; shift 3
K="timeout -s KILL"
mkdir -p gpurun_out
SMALL="python bench.py --chains 16384 --steps 3 --warmup 44 --no-e2e --no-cpu-baseline --no-breakdown $@"
$K 200 $SMALL > gpurun_out/${TAG}_plain.log 2>&1 && \
$K 600 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s $SKIP -c 1 -o gpurun_out/${TAG} -f $SMALL > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
