#!/bin/bash
# Round 2, GPU call 15 (1 GPU): whole GPU suite at HEAD (new: long chains vs oracle, flags variants), halo-warp
# loop A/B, then the round's profile set (bench line + launch list + ncu --set full) for configs 1-3.
mkdir -p gpurun_out
O=gpurun_out/r2c15
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
K=d4r3w12p5
timeout 600 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 $K SFB200_HALO_SKIP=1:$K SFB200_SCHED=halving:$K > ${O}_sweep1.txt 2>&1
grep -A6 medians ${O}_sweep1.txt; grep -i "differ\|fail" ${O}_sweep1.txt | head
bash scripts/profile_round.sh r02b > ${O}_profile_round.txt 2>&1
tail -12 ${O}_profile_round.txt
