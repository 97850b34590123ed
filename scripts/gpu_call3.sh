#!/bin/bash
mkdir -p gpurun_out
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c3_pytest.txt 2>&1
tail -4 gpurun_out/c3_pytest.txt
for c in 1 2 3; do
  timeout 900 python bench.py --config $c --steps 10 --warmup 3 > gpurun_out/r01b_config${c}_bench.json 2> gpurun_out/r01b_config${c}_bench.err
  tail -c 1500 gpurun_out/r01b_config${c}_bench.json; echo
done
free -g | head -2
timeout 600 python bench.py --config 4 --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01b_config4_bench.json 2> gpurun_out/r01b_config4_bench.err
tail -c 1500 gpurun_out/r01b_config4_bench.json; tail -3 gpurun_out/r01b_config4_bench.err
