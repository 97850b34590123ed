#!/bin/bash
# Round 2, GPU call 29 (1 GPU): work-list order -- edge tiles first vs edge rows first (index order kept).
mkdir -p gpurun_out
O=gpurun_out/r2c29
timeout 600 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 4 d4r3w12p5 SFB200_SCHED=rows:d4r3w12p5 SFB200_SCHED=halving:d4r3w12p5 > ${O}_sweep1.txt 2>&1
grep -A4 medians ${O}_sweep1.txt; grep -i "differ\|fail\|lower" ${O}_sweep1.txt | head -4
