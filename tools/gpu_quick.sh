#!/bin/bash
# quick A/B loop on the B200 box: selected tests + short benches. usage: tools/gpu_quick.sh <tag> ["pytest -k expr"] [extra bench args]
TAG=${1:-q}
KEXPR=${2:-"inc or round2 or parity"}
K="timeout -s KILL"
mkdir -p gpurun_out
$K 600 python -m pytest tests -m gpu -x -q -k "$KEXPR" 2>&1 | tail -6 | tee gpurun_out/${TAG}_pytest.log
B="python bench.py --steps 20 --warmup 40 --no-e2e --no-cpu-baseline"
$K 300 $B > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/${TAG}_bench.json").read().strip().splitlines()[-1])
    print("value %.4g  ms/step %.3f" % (d["value"], d["ms_per_step"]))
    print({k: round(v, 3) for k, v in d["kernel_ms"].items()})
except Exception as e:
    print("bench parse failed", e); print(open("gpurun_out/${TAG}_bench.err").read()[-2000:])
PY
