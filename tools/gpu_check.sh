#!/bin/bash
# Runs on the B200 box (under gpurun): GPU parity tests, smoke, full default bench, reference arm,
# ncu launch list and `--set full` captures of the step kernels. Outputs under gpurun_out/.
# usage: tools/gpu_check.sh <tag> [quick]
TAG=${1:-r02}
MODE=${2:-full}
K="timeout -s KILL"
mkdir -p gpurun_out
$K 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/${TAG}_pytest.log
$K 120 python __graft_entry__.py smoke 2>&1 | tail -2 | tee gpurun_out/${TAG}_smoke.log
$K 600 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json; tail -5 gpurun_out/${TAG}_bench.err
if [ "$MODE" = "full" ]; then
$K 400 python bench.py --impl reference --steps 4 --warmup 1 > gpurun_out/${TAG}_bench_ref.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ref.json
$K 200 python bench.py --workload ube4b_potts_poe_4k --no-cpu-baseline > gpurun_out/${TAG}_bench_ube4b.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_ube4b.json
$K 300 python bench.py --workload gfp_paper_pas10 --no-cpu-baseline --steps 5 > gpurun_out/${TAG}_bench_pas10.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_pas10.json
$K 300 python bench.py --workload pabp_readme_128 --steps 100 --warmup 5 > gpurun_out/${TAG}_bench_pabp128.json 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_bench_pabp128.json
fi
SMALL="python bench.py --chains 16384 --steps 3 --warmup 44 --no-e2e --no-cpu-baseline --no-breakdown"
$K 200 $SMALL > gpurun_out/${TAG}_plain.log 2>&1 && \
$K 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 380 -c 110 --csv --log-file gpurun_out/${TAG}_launches.csv $SMALL > gpurun_out/${TAG}_ncu_l.log 2>&1
# steady state (44 warm-up iterations, 7 matching kernels per delta iteration): skip ~41 iterations, capture two
$K 200 $SMALL > gpurun_out/${TAG}_plain2.log 2>&1 && \
$K 600 ncu --set full --clock-control none --import-source on -k regex:"cnn_forward_inc_kernel|cnn_inc_merge_kernel|cnn_backward_delta_kernel|cnn_delta_record_kernel|pas_propose|pas_reverse_accept|cnn_fit_kernel" -s 290 -c 14 -o gpurun_out/${TAG}_full -f $SMALL > gpurun_out/${TAG}_ncu_f.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_f.log
if [ "$MODE" = "full" ]; then
$K 500 python tools/bench_potts_full.py 64 128 238 512 1024 > gpurun_out/${TAG}_potts_full_sweep.jsonl 2>> gpurun_out/${TAG}_bench.err; cat gpurun_out/${TAG}_potts_full_sweep.jsonl
PD="python tools/bench_potts_full.py 238"
$K 200 $PD > gpurun_out/${TAG}_plain3.log 2>&1 && \
$K 300 ncu --set full --clock-control none --import-source on -k regex:potts_dense_tc_kernel -s 6 -c 1 -o gpurun_out/${TAG}_potts_dense -f $PD > gpurun_out/${TAG}_ncu_p.log 2>&1
tail -1 gpurun_out/${TAG}_ncu_p.log
fi
