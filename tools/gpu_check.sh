#!/bin/bash
# Runs on the B200 box (under gpurun): GPU parity tests, smoke, full default bench, reference arm,
# ncu launch list and one `--set full` capture of the top kernels. Outputs under gpurun_out/.
# usage: tools/gpu_check.sh <tag> [kernel-regex for the full capture]
TAG=${1:-r01}
KRE=${2:-cnn_forward_tc2_kernel}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/${TAG}_pytest.log
python __graft_entry__.py smoke 2>&1 | tail -3 | tee gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
SMALL="python bench.py --chains 8192 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline"
$SMALL > gpurun_out/${TAG}_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu_l.log 2>&1
$SMALL > gpurun_out/${TAG}_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"$KRE" -s 4 -c 2 -o gpurun_out/${TAG}_full -f $SMALL > gpurun_out/${TAG}_ncu_f.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_f.log
