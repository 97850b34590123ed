#!/bin/bash
# Round 2, GPU call 12 (8 GPUs): bench --gpus 8 (weak config 1 + verify vs one GPU + configs[4] strong block with the
# one-GPU leg from the same box), the one-GPU weak baseline of the same box, host DMA ceiling with 8 ranks.
mkdir -p gpurun_out
O=gpurun_out/r2c12
nvidia-smi topo -m > ${O}_topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 20 ) > ${O}_bench_8gpu.json 2> ${O}_bench_8gpu.err
tail -c 4500 ${O}_bench_8gpu.json; tail -5 ${O}_bench_8gpu.err
( time timeout 300 python bench.py --gpus 1 --steps 20 --no-e2e --no-cpu-baseline --no-strong ) > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err
tail -c 1800 ${O}_bench_1gpu.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 scripts/host_dma_ceiling.py > ${O}_dma_8gpu.json 2> ${O}_dma_8gpu.err; cat ${O}_dma_8gpu.json
