mkdir -p gpurun_out
for v in 0 7 9; do
BVCF_SITES_VAR=$v python bench.py --config c3 --steps 3 --warmup 3 --no-bgzf --no-cpu-baseline --e2e-lines 100000 > gpurun_out/var_c3_v$v.json 2> gpurun_out/var_c3_v$v.err
done
python - <<'P'
import json
for f in ("var_c3_v0","var_c3_v7","var_c3_v9"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_checked"))
    except Exception as e: print(f, "ERR", e)
P
