#!/bin/bash
# One B200: bench lines for configs 1-3, the ncu launch lists of the same commands, and one
# `ncu --set full` capture of the dominant kernel per config.  Outputs land in gpurun_out/.
TAG=${1:-r01}
for c in 1 2 3; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/bench_${TAG}_config$c.json 2> gpurun_out/bench_${TAG}_config$c.err
  tail -c 2500 gpurun_out/bench_${TAG}_config$c.json
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file gpurun_out/launches_${TAG}_config$c.csv \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sf_ -s 3 -c 1 \
      -o gpurun_out/prof_${TAG}_config$c -f \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
done
ls -la gpurun_out | tail -20
