#!/bin/bash
mkdir -p gpurun_out
export SFB200_MAX_DEPTH=4 SFB200_ROWS=4 SFB200_WARPS=8 SFB200_PREFETCH=5
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sf_stream -s 3 -c 1 \
   -o gpurun_out/s4_c1_p5_full -f \
   python bench.py --config 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/s4_c1_p5_ncu.log 2>&1
ls -la gpurun_out/
