#!/bin/bash
# Round 2, GPU call 39 (1 GPU): 1-warp CTAs for the 2-D chain (chunk length, ring depth); small CTAs for the 3-D chain.
mkdir -p gpurun_out
O=gpurun_out/r2c39
timeout 900 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 d8v4w2p5 d8v4w1p5 SFB200_CHUNK=224:d8v4w1p5 SFB200_CHUNK=288:d8v4w1p5 SFB200_CHUNK=352:d8v4w1p5 SFB200_CHUNK=416:d8v4w1p5 SFB200_CHUNK=640:d8v4w1p5 SFB200_CHUNK=1024:d8v4w1p5 d8v4w1p3 d8v4w1p2 SFB200_SPLITLOOP=0:d8v4w1p5 > ${O}_sweep3.txt 2>&1
grep -A12 medians ${O}_sweep3.txt; grep -i "differ\|fail\|lower" ${O}_sweep3.txt | head -3
timeout 900 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 d4r3w12p5 d4r3w6p5 SFB200_PERSISTENT=0:d4r3w6p5 SFB200_PERSISTENT=0,SFB200_CHUNK=96:d4r3w6p5 SFB200_PERSISTENT=0,SFB200_CHUNK=160:d4r3w6p5 SFB200_PERSISTENT=0,SFB200_CHUNK=256:d4r3w6p5 SFB200_PERSISTENT=0,SFB200_CHUNK=160:d4r4w4p5 > ${O}_sweep1.txt 2>&1
grep -A8 medians ${O}_sweep1.txt; grep -i "differ\|fail\|lower" ${O}_sweep1.txt | head -3
