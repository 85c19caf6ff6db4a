#!/bin/bash
# GPU box helper: parity tests + a short C2 bench, compact output.  usage: bash tools/gpu_check.sh [tag] [lines] [pytest-args]
TAG=${1:-chk}
LINES=${2:-300000}
PYARGS=${3:-tests}
mkdir -p gpurun_out
timeout 1500 python -m pytest $PYARGS -m gpu -x -q 2>&1 | tail -25 > gpurun_out/pytest_$TAG.log
for i in 1 2; do
  timeout 600 python bench.py --lines $LINES --steps 3 --warmup 3 --e2e-lines 20000 --no-cpu-baseline > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/plain_$TAG.log").read().strip().splitlines()[-1])
    print(round(d["value"]/1e6,1), {k: round(v,4) for k,v in d["kernel_ms_per_step"].items()}, round(d["roofline"]["frac"],4))
except Exception as e:
    print("bench failed", e); print(open("gpurun_out/plain_$TAG.err").read()[-1500:])
PY
done
echo "PYTEST:"; cat gpurun_out/pytest_$TAG.log
