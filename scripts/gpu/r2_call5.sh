#!/bin/bash
# Round 2, GPU call 5 (1 GPU): full parity suite, the default bench line (verify, e2e, strong block, CPU
# baseline), reference arm, host DMA ceiling.
mkdir -p gpurun_out
O=gpurun_out/r2c5
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
( time timeout 900 python bench.py --steps 20 ) > ${O}_bench_default.json 2> ${O}_bench_default.err
tail -c 4000 ${O}_bench_default.json; tail -5 ${O}_bench_default.err
( time timeout 600 python bench.py --impl reference --steps 20 --warmup 5 ) > ${O}_bench_reference.json 2> ${O}_bench_reference.err
tail -c 1500 ${O}_bench_reference.json
timeout 300 python scripts/host_dma_ceiling.py > ${O}_dma_1gpu.json 2> ${O}_dma_1gpu.err; cat ${O}_dma_1gpu.json
for cfg in 2 3; do
  timeout 600 python bench.py --config $cfg --steps 20 --no-cpu-baseline > ${O}_bench_cfg$cfg.json 2> ${O}_bench_cfg$cfg.err
  tail -c 2500 ${O}_bench_cfg$cfg.json; tail -3 ${O}_bench_cfg$cfg.err
done
