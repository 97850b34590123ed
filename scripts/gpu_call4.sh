#!/bin/bash
# Session 8, call 1: GPU parity suite on HEAD, then plan-variant sweeps (higher-occupancy geometries).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/c4_smi.txt
( time timeout 1100 python -m pytest tests -m gpu -x -q ) > gpurun_out/c4_pytest.txt 2>&1
tail -5 gpurun_out/c4_pytest.txt
timeout 400 python scripts/sweep_variants.py --config 1 d4r3w12p5 d4r2w16p5 d4r2w16p4 d4r2w16p3 d4r2w16k32p3 d4r2w16p5s \
    d4r2w12p5 d5r2w12p5 d4r3w10p5 d4r3w12p4 d4r3w12p6 d3r2w16p5 d4r3w12p5 > gpurun_out/c4_sweep1.txt 2>&1
cat gpurun_out/c4_sweep1.txt
timeout 400 python scripts/sweep_variants.py --config 3 d8v4w8p5 d8v4w8p4 d8v4w8p6 d8v6w8p5 d8v2w12p5 d8v8w8p5 d16v4w8p5 \
    d8v4w8p5 > gpurun_out/c4_sweep3.txt 2>&1
cat gpurun_out/c4_sweep3.txt
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/c4_bench_default.json 2> gpurun_out/c4_bench_default.err
tail -c 1200 gpurun_out/c4_bench_default.json
