#!/bin/bash
# round-2 late pass: final C2 line, names scatter A/B, C3 stage times + one `ncu --set full` of compose / copy-out on C3
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/final_c2.json 2> gpurun_out/final_c2.err
BVCF_NAMES_ITER=1 python bench.py --steps 5 --warmup 3 --no-bgzf --no-cpu-baseline --e2e-lines 50000 > gpurun_out/iter_c2.json 2> gpurun_out/iter_c2.err
python bench.py --config c3 --steps 3 --warmup 3 --no-bgzf --no-cpu-baseline > gpurun_out/final_c3.json 2> gpurun_out/final_c3.err
CMD="python bench.py --config c3 --lines 4000000 --steps 1 --warmup 3 --e2e-lines 20000 --no-cpu-baseline --no-bgzf"
ncu --set full --clock-control none --import-source on -k regex:"bvcf_compose|bvcf_copyout" -s 6 -c 2 -o gpurun_out/prof_c3b $CMD > gpurun_out/ncu_c3b.log 2>&1
tail -2 gpurun_out/ncu_c3b.log
python - <<'P'
import json
for f in ("final_c2","iter_c2","final_c3"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_checked"))
    except Exception as e: print(f, "ERR", e)
P
