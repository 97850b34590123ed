#!/bin/bash
# Round 2, GPU call 35 (1 GPU): halo-warp skip under the final defaults; persistent CTAs / chunk length for the 2-D chain.
mkdir -p gpurun_out
O=gpurun_out/r2c35
timeout 600 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 4 d4r3w12p5 SFB200_HALO_SKIP=1:d4r3w12p5 > ${O}_sweep1.txt 2>&1
grep -A3 medians ${O}_sweep1.txt; grep -i "differ\|fail\|lower" ${O}_sweep1.txt | head -3
timeout 600 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 d8v4w2p5 SFB200_PERSISTENT=1:d8v4w2p5 SFB200_PERSISTENT=1,SFB200_SCHED=halving:d8v4w2p5 SFB200_CHUNK=964:d8v4w2p5 SFB200_CHUNK=1130:d8v4w2p5 > ${O}_sweep3.txt 2>&1
grep -A6 medians ${O}_sweep3.txt; grep -i "differ\|fail\|lower" ${O}_sweep3.txt | head -3
