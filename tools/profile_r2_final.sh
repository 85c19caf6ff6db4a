#!/bin/bash
# End-of-round evidence on one B200: the four BASELINE configs through bench.py, then the ncu recipe of
# /opt/skills/guides/B200_PROFILING.md on the 300,000-variant C2 shard (launch list + one --set full capture of a
# whole resident step) and on a 4 M-line C3 shard (compose / copy-out / scan of the sites-only path).
mkdir -p gpurun_out
python bench.py --steps 5 --warmup 3 > gpurun_out/fin_c2.json 2> gpurun_out/fin_c2.err
for c in c3 c4 c5; do
  python bench.py --config $c --steps 3 --warmup 3 --no-bgzf > gpurun_out/fin_$c.json 2> gpurun_out/fin_$c.err
done
python - <<'P'
import json
for f in ("fin_c2","fin_c3","fin_c4","fin_c5"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, round(d["ms_per_step"],3), {k:round(v,3) for k,v in d["kernel_ms_per_step"].items()}, d.get("parity_checked"),
              "e2e", round(d["e2e"]["value"]/1e6,2), "pipe", round(d.get("roofline_pipeline",{}).get("frac",0),4), "cpu", d.get("cpu_baseline",{}).get("value"))
        if d.get("e2e_bgzf"): print("   bgzf", {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["e2e_bgzf"].items() if not isinstance(v,(dict,str))})
    except Exception as e: print(f, "ERR", e)
P
CMD="python bench.py --lines 300000 --steps 2 --warmup 3 --e2e-lines 20000 --no-cpu-baseline --no-bgzf"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bvcf_ -s 60 -c 20 --csv --log-file gpurun_out/launches_fin.csv $CMD > gpurun_out/ncu_l_fin.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvcf_ -s 60 -c 20 -o gpurun_out/prof_fin $CMD > gpurun_out/ncu_f_fin.log 2>&1
tail -1 gpurun_out/ncu_f_fin.log
CMD="python bench.py --config c3 --lines 4000000 --steps 1 --warmup 3 --e2e-lines 20000 --no-cpu-baseline --no-bgzf"
ncu --set full --clock-control none --import-source on -k regex:"bvcf_compose|bvcf_copyout|bvcf_scan_genotype|bvcf_compact" -s 12 -c 4 -o gpurun_out/prof_fin_c3 $CMD > gpurun_out/ncu_f_fin_c3.log 2>&1
tail -1 gpurun_out/ncu_f_fin_c3.log
