mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tile.py tests/test_gpu_parity.py -x -q -m gpu -k "sites or slow_path or parity or many or long_info" > gpurun_out/sites_tests.log 2>&1
tail -5 gpurun_out/sites_tests.log
python bench.py --config c3 --steps 3 --warmup 3 --no-bgzf --no-cpu-baseline > gpurun_out/sites_c3.json 2> gpurun_out/sites_c3.err
python - <<'P'
import json
for f in ("sites_c3",):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_checked"), d["e2e"]["value"])
    except Exception as e: print(f, "ERR", e)
P
tail -3 gpurun_out/sites_c3.err
