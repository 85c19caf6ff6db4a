#!/bin/bash
# ncu recipe of /opt/skills/guides/B200_PROFILING.md for the round-2 pipeline: plain bench first, then a launch list
# and one full capture of a whole pipeline step (the resident 300,000-variant run), plus the inflate kernel.
# usage (on the GPU box, via gpurun): bash tools/profile_r2.sh <tag> [lines]
TAG=${1:-r2}
LINES=${2:-300000}
mkdir -p gpurun_out
CMD="python bench.py --lines $LINES --steps 2 --warmup 3 --e2e-lines 20000 --no-cpu-baseline --bgzf-mb 256"
$CMD > gpurun_out/plain_$TAG.log 2> gpurun_out/plain_$TAG.err || { tail -5 gpurun_out/plain_$TAG.err; exit 1; }
# warm-up runs 1-3 and timed runs 4-5 each launch 20 kernels: skip the first three runs
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:bvcf_ -s 60 -c 20 --csv --log-file gpurun_out/launches_$TAG.csv $CMD > gpurun_out/ncu_l_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvcf_ -s 60 -c 20 -o gpurun_out/prof_$TAG $CMD > gpurun_out/ncu_f_$TAG.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:bvcf_inflate -s 3 -c 1 -o gpurun_out/prof_${TAG}_inflate $CMD > gpurun_out/ncu_i_$TAG.log 2>&1
tail -1 gpurun_out/plain_$TAG.log | cut -c1-600
tail -2 gpurun_out/ncu_f_$TAG.log
