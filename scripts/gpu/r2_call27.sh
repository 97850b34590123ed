#!/bin/bash
# Round 2, GPU call 27 (1 GPU): HEAD (balanced sums for both float types): whole GPU suite, profile set r02d.
mkdir -p gpurun_out
O=gpurun_out/r2c27
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
bash scripts/profile_round.sh r02d > ${O}_profile_round.txt 2>&1
grep -E "gpu__time|traffic /" ${O}_profile_round.txt
( time timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 ) > ${O}_bench_driver_cmd.json 2> ${O}_bench_driver_cmd.err
tail -c 1500 ${O}_bench_driver_cmd.json
