#!/bin/bash
# Round 2, GPU call 25 (1 GPU): new defaults (per-CTA boundary test with the longest-first list, balanced float64
# sums): whole GPU suite, A/B of the variants that no longer spill, profile set r02c for configs 1-3.
mkdir -p gpurun_out
O=gpurun_out/r2c25
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
K=d4r3w12p5
timeout 600 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 $K SFB200_BC_MODE=thread:$K SFB200_HALO_SKIP=1:$K SFB200_ST64=1:$K ${K}h > ${O}_sweep1.txt 2>&1
grep -A6 medians ${O}_sweep1.txt; grep -i "differ\|fail\|lower" ${O}_sweep1.txt | head -5
bash scripts/profile_round.sh r02c > ${O}_profile_round.txt 2>&1
grep -E "gpu__time|traffic /" ${O}_profile_round.txt
