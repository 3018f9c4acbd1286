#!/bin/bash
# ncu capture of the unprojection kernel (device-resident rasters, config 2) -> gpurun_out/prof_unproj_<tag>.ncu-rep
tag=${1:-x}
mkdir -p gpurun_out
python tools/unproject_timing.py c2 > gpurun_out/unproj_plain_$tag.log 2>&1 && cat gpurun_out/unproj_plain_$tag.log &&
timeout 300 ncu --set full --clock-control none --import-source on -k regex:unproject_fused -s 2 -c 1 -f \
  -o gpurun_out/prof_unproj_$tag python tools/unproject_timing.py c2 > gpurun_out/ncu_unproj_$tag.log 2>&1
tail -2 gpurun_out/ncu_unproj_$tag.log
