#!/bin/bash
# Round 2, GPU call 24 (1 GPU): the remaining code-shape switches under the round-2 defaults (L2 hints, boundary
# fix-up mode, software pipelining, ring depth, re-association), configs 1 and 3.
mkdir -p gpurun_out
O=gpurun_out/r2c24
timeout 900 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 d4r3w12p5 SFB200_L2HINT=0:d4r3w12p5 SFB200_L2HINT=1:d4r3w12p5 SFB200_L2HINT=2:d4r3w12p5 SFB200_BC_MODE=cta:d4r3w12p5 SFB200_PIPELINE=0:d4r3w12p5 d4r3w12p3 d4r3w12p2 SFB200_REASSOCIATE=0:d4r3w12p5 SFB200_REASSOCIATE=2:d4r3w12p5 > ${O}_sweep1.txt 2>&1
grep -A12 medians ${O}_sweep1.txt; grep -i "differ\|fail\|lower" ${O}_sweep1.txt | head
timeout 900 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 d8v4w2p5 SFB200_L2HINT=0:d8v4w2p5 SFB200_REASSOCIATE=3:d8v4w2p5 SFB200_PIPELINE=0:d8v4w2p5 d8v4w2p3 SFB200_BC_MODE=thread:d8v4w2p5 > ${O}_sweep3.txt 2>&1
grep -A8 medians ${O}_sweep3.txt; grep -i "differ\|fail\|lower" ${O}_sweep3.txt | head
