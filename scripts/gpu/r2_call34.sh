#!/bin/bash
# Round 2, GPU call 34 (1 GPU): the new generator-switch parity tests.
mkdir -p gpurun_out
O=gpurun_out/r2c34
( time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "generator_switches or pair_sync" ) > ${O}_pytest.txt 2>&1
tail -6 ${O}_pytest.txt
