#!/bin/bash
# Round-end check on one B200: parity tests, smoke, the default bench line, bench lines of the
# other configs, the ncu launch list and one full capture of the step kernel (steady state).
# Usage (from the repo root, under gpurun):  bash tools/final_check.sh <tag>
tag=${1:-x}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -x -q > $out/pytest_gpu_$tag.log 2>&1; tail -2 $out/pytest_gpu_$tag.log
python __graft_entry__.py smoke > $out/smoke_$tag.log 2>&1; tail -1 $out/smoke_$tag.log
python bench.py > $out/bench_${tag}_c2.json 2> $out/bench_${tag}_c2.err
for c in c3 c4 c5; do
  python bench.py --no-e2e --no-cpu --config $c --steps 10 --warmup 3 > $out/bench_${tag}_$c.json 2> $out/bench_${tag}_$c.err
done
python tools/bench_brief.py $out/bench_${tag}_c*.json
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches_$tag.csv \
  python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 > $out/ncu1_$tag.log 2>&1
timeout 120 ncu --set full --clock-control none --import-source on -k regex:lloyd_step -s 32 -c 2 -f -o $out/prof_step_$tag \
  python bench.py --no-e2e --no-cpu --steps 3 --warmup 3 > $out/ncu2_$tag.log 2>&1
tail -1 $out/ncu2_$tag.log
