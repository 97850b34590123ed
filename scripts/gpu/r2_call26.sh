#!/bin/bash
# Round 2, GPU call 26 (1 GPU): second-order sweep under the new defaults, configs 1 and 3.
mkdir -p gpurun_out
O=gpurun_out/r2c26
timeout 900 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 d4r3w12p5 SFB200_UNROLL=12,SFB200_UNROLL_MULT=2:d4r3w12p5 d4r4w8p5 d4r4w8k16p5 SFB200_SCHED=halving:d4r3w12p5 SFB200_L2HINT=1:d4r3w12p5 SFB200_REASSOCIATE=2:d4r3w12p5 > ${O}_sweep1.txt 2>&1
grep -A9 medians ${O}_sweep1.txt; grep -i "differ\|fail\|lower" ${O}_sweep1.txt | head -4
timeout 900 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 d8v4w2p5 SFB200_SPLITBAR=1:d8v4w2p5 d8v4w4p5 SFB200_UNROLL=12,SFB200_UNROLL_MULT=2:d8v4w2p5 SFB200_BC_MODE=thread:d8v4w2p5 SFB200_L2HINT=0:d8v4w2p5 d8v2w2p5 d8v2w4p5 > ${O}_sweep3.txt 2>&1
grep -A10 medians ${O}_sweep3.txt; grep -i "differ\|fail\|lower" ${O}_sweep3.txt | head -4
