#!/bin/bash
# Round 2, GPU call 38 (1 GPU): CTA width of the 2-D chain under the short-chunk default.
mkdir -p gpurun_out
O=gpurun_out/r2c38
timeout 900 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 d8v4w2p5 d8v4w1p5 d8v4w4p5 SFB200_CHUNK=416:d8v4w2p5 SFB200_CHUNK=512:d8v4w2p5 SFB200_CHUNK=380:d8v4w2p5 > ${O}_sweep3.txt 2>&1
grep -A7 medians ${O}_sweep3.txt; grep -i "differ\|fail\|lower" ${O}_sweep3.txt | head -3
