#!/bin/bash
mkdir -p gpurun_out
export C3_LINES=4000000
python tools/perf_configs.py C3 > gpurun_out/plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"bvcf_rows|bvcf_scan" -s 4 -c 3 -o gpurun_out/prof_c3 python tools/perf_configs.py C3 > gpurun_out/ncu_c3.log 2>&1
tail -2 gpurun_out/plain_c3.log | cut -c1-400
