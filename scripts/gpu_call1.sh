#!/bin/bash
# round-1 session-4 GPU call: regression tests, then the synchronisation / prefetch sweeps
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/c1_gpu.txt 2>&1
( time timeout 1200 python -m pytest tests -m gpu -x -q ) > gpurun_out/c1_pytest.txt 2>&1
tail -5 gpurun_out/c1_pytest.txt
timeout 400 python scripts/sweep_variants.py --config 1 d4r4w8 d4r4w8p3 d4r4w8p5 d4r4w8s d4r4w8p3s d4r4w8p5s d4r4w8k32p5s d3r4w12p5s d4r3w12 d4r3w12p5 d4r3w12p5s d2r4w16s > gpurun_out/c1_sweep1.txt 2>&1
cat gpurun_out/c1_sweep1.txt | grep -v Warning
timeout 400 python scripts/sweep_variants.py --config 3 d8v4w8 d8v4w8p3 d8v4w8p5 d8v4w8s d8v4w8p5s d8v2w16p5s d8v2w16p5 d16v2w16p5s d6v4w8p5s d4v4w16p5s > gpurun_out/c1_sweep3.txt 2>&1
cat gpurun_out/c1_sweep3.txt | grep -v Warning
timeout 200 python scripts/sweep_variants.py --config 2 d4r2 d4r2p3 d4r2p5 d4r4 d4r4p5 > gpurun_out/c1_sweep2.txt 2>&1
cat gpurun_out/c1_sweep2.txt | grep -v Warning
