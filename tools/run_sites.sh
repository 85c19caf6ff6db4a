mkdir -p gpurun_out
python bench.py --config c5 --steps 3 --warmup 3 --no-bgzf --no-cpu-baseline --e2e-lines 100000 > gpurun_out/dz_c5.json 2> gpurun_out/dz_c5.err
python - <<'P'
import json
for f in ("dz_c5",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_checked"))
    except Exception as e: print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
P
timeout 900 python -m pytest tests -x -q -m gpu -k "dosage or locus or slow_path or c5 or arrow or feather" 2>&1 | tail -3
