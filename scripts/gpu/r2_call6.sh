#!/bin/bash
# Round 2, GPU call 6 (1 GPU): default bench line again (strong block fixed), new parity cases, DMA ceiling incl. registered memory.
mkdir -p gpurun_out
O=gpurun_out/r2c6
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
( time timeout 900 python bench.py --steps 20 ) > ${O}_bench_default.json 2> ${O}_bench_default.err
tail -c 5000 ${O}_bench_default.json; tail -5 ${O}_bench_default.err
timeout 300 python scripts/host_dma_ceiling.py > ${O}_dma_1gpu.json 2> ${O}_dma_1gpu.err; cat ${O}_dma_1gpu.json; tail -3 ${O}_dma_1gpu.err
