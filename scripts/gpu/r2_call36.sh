#!/bin/bash
# Round 2, GPU call 36 (1 GPU): rows per CTA (chunk length) of the 2-D chain's (tile, chunk) grid.
mkdir -p gpurun_out
O=gpurun_out/r2c36
timeout 900 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 d8v4w2p5 SFB200_CHUNK=160:d8v4w2p5 SFB200_CHUNK=224:d8v4w2p5 SFB200_CHUNK=288:d8v4w2p5 SFB200_CHUNK=352:d8v4w2p5 SFB200_CHUNK=416:d8v4w2p5 SFB200_CHUNK=482:d8v4w2p5 > ${O}_sweep3b.txt 2>&1
grep -A11 medians ${O}_sweep3b.txt; grep -i "differ\|fail\|lower" ${O}_sweep3b.txt | head -3
