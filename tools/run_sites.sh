mkdir -p gpurun_out
LINES=3000000 BVCF_TRACE=1 timeout 600 python tools/e2e_probe.py 2>&1 | grep -v "^\[bvcf\]" | tail -5
python bench.py --config c3 --steps 3 --warmup 3 --no-bgzf --no-cpu-baseline > gpurun_out/dg_c3.json 2> gpurun_out/dg_c3.err
python - <<'P'
import json
d=json.loads(open("gpurun_out/dg_c3.json").read().strip().splitlines()[-1])
print(round(d["ms_per_step"],3), d["e2e"])
P
timeout 900 python -m pytest tests -x -q -m gpu -k "diag or cli or golden or fuzz" 2>&1 | tail -3
