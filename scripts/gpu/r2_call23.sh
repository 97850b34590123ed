#!/bin/bash
# Round 2, GPU call 23 (1 GPU): "halves" synchronisation -- parity of the new code shape, then A/B timing.
mkdir -p gpurun_out
O=gpurun_out/r2c23
( time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "pair_sync" ) > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
K=d4r3w12p5
timeout 900 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 $K ${K}h d4r4w8p5 d4r4w8p5h d4r4w8p2h d4r3w12p2h d4r2w16p5h > ${O}_sweep1.txt 2>&1
grep -A10 medians ${O}_sweep1.txt; grep -i "differ\|fail\|lower" ${O}_sweep1.txt | head
