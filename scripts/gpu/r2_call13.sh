#!/bin/bash
# Round 2, GPU call 13 (1 GPU): "flags" synchronisation (per-field mbarriers instead of the CTA barrier), split
# CTA barrier, 64-bit shared stores -- parity of the new code shapes, then A/B timing round-robin on configs 1, 3.
mkdir -p gpurun_out
O=gpurun_out/r2c13
( time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "pair_sync or flags" ) > ${O}_pytest.txt 2>&1
tail -6 ${O}_pytest.txt
K=d4r3w12p5
timeout 900 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 $K SFB200_ST64=1:$K SFB200_SPLITBAR=1:$K ${K}f SFB200_ST64=1:${K}f d4r4w8p5 d4r4w8p5f d4r4w8p2f d4r3w12p4x SFB200_ST64=1:d4r3w12p4x d4r4w8p4x > ${O}_sweep1.txt 2>&1
grep -A20 medians ${O}_sweep1.txt; grep -i "differ\|fail" ${O}_sweep1.txt | head
K3=d8v4w2p5
timeout 900 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 $K3 SFB200_SPLITBAR=1:$K3 ${K3}f d8v4w4p5 d8v4w4p5f d8v4w8p5f d8v4w1p5f > ${O}_sweep3.txt 2>&1
grep -A20 medians ${O}_sweep3.txt; grep -i "differ\|fail" ${O}_sweep3.txt | head
