#!/bin/bash
# ncu capture of steady-state Lloyd iterations of a config -> gpurun_out/prof_step_<cfg>_<tag>.ncu-rep
tag=${1:-x}; cfg=${2:-c2}; skip=${3:-47}
mkdir -p gpurun_out
args="--no-e2e --no-cpu --no-stream-all --no-bruteforce --config $cfg --steps 3 --warmup 3"
python bench.py $args > gpurun_out/plain_${cfg}_$tag.log 2>&1 && python tools/bench_brief.py gpurun_out/plain_${cfg}_$tag.log &&
timeout 300 ncu --set full --clock-control none --cache-control none --import-source on -k regex:lloyd_step -s $skip -c 2 -f \
  -o gpurun_out/prof_step_${cfg}_$tag python bench.py $args > gpurun_out/ncu_step_${cfg}_$tag.log 2>&1
tail -1 gpurun_out/ncu_step_${cfg}_$tag.log
