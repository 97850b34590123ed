#!/bin/bash
# Session 8 wrap-up call: full GPU parity suite, default bench line (with traffic), reference arm, bench variants.
mkdir -p gpurun_out
( time timeout 1100 python -m pytest tests -m gpu -q ) > gpurun_out/c9_pytest.txt 2>&1
tail -5 gpurun_out/c9_pytest.txt
timeout 400 python bench.py > gpurun_out/c9_bench_default.json 2> gpurun_out/c9_bench_default.err
tail -c 1500 gpurun_out/c9_bench_default.json; echo
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/c9_bench_reference.json 2> gpurun_out/c9_bench_reference.err
cat gpurun_out/c9_bench_reference.json; tail -3 gpurun_out/c9_bench_reference.err
timeout 300 python bench.py --config 2 --variant jki --steps 20 --no-cpu-baseline > gpurun_out/r01c_config2_jki_bench.json 2> gpurun_out/c9_jki.err
tail -c 700 gpurun_out/r01c_config2_jki_bench.json; echo
timeout 400 python bench.py --config 3 --variant w1d --steps 10 --no-cpu-baseline > gpurun_out/r01c_config3_w1d_bench.json 2> gpurun_out/c9_w1d.err
tail -c 700 gpurun_out/r01c_config3_w1d_bench.json; echo
