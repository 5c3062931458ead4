#!/bin/bash
# Multi-GPU records on one B200 box (under `gpurun --gpus 8`): the log_every path under NCCL with a global population, and the
# strong-scaling series of BASELINE.json configs[2] (65,536 chains in total over 1/2/4/8 GPUs).  usage: tools/gpu_multi.sh <tag>
TAG=${1:-r02}
K="timeout -s KILL 600"
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
mkdir -p gpurun_out
$K $TR --nproc-per-node 8 --master-port 29521 bench.py --gpus 8 --steps 20 --warmup 5 --log-every 5 --global-population --no-cpu-baseline \
    > gpurun_out/${TAG}_8gpu_log5_global.json 2> gpurun_out/${TAG}_multi.err; echo "log5 rc=$?"
for N in 8 4 2; do
  $K $TR --nproc-per-node $N --master-port $((29530 + N)) bench.py --gpus $N --strong --steps 20 --warmup 5 --no-cpu-baseline --no-breakdown \
      > gpurun_out/${TAG}_strong_${N}gpu.json 2>> gpurun_out/${TAG}_multi.err; echo "strong $N rc=$?"
done
$K python bench.py --gpus 1 --strong --steps 20 --warmup 5 --no-cpu-baseline --no-breakdown > gpurun_out/${TAG}_strong_1gpu.json 2>> gpurun_out/${TAG}_multi.err; echo "strong 1 rc=$?"
$K $TR --nproc-per-node 8 --master-port 29541 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline --no-breakdown \
    > gpurun_out/${TAG}_weak_8gpu.json 2>> gpurun_out/${TAG}_multi.err; echo "weak 8 rc=$?"
# BASELINE.json configs[4] (Potts-only sweep) on all GPUs: every rank the same sweep on its own chains, times = max over ranks
$K $TR --nproc-per-node 8 --master-port 29551 tools/bench_potts_full.py 238 > gpurun_out/${TAG}_potts_sweep_8gpu.jsonl 2>> gpurun_out/${TAG}_multi.err; echo "potts sweep 8 rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/${TAG}_*gpu*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        e = d.get("e2e") or {}
        print(f.split("/")[-1], "n_gpus", d["n_gpus"], "value %.4g" % d["value"], "ms %.3f" % d["ms_per_step"], "e2e %.4g" % e.get("value", 0),
              "reports", e.get("log_report_device_ms"))
    except Exception as ex:
        print(f, "parse failed", ex)
PY
tail -3 gpurun_out/${TAG}_multi.err
