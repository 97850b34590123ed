#!/bin/bash
# Round 2, GPU call 11 (4 GPUs): bench --gpus 4 (weak config 1, verify vs one GPU, configs[4] strong block with the
# one-GPU leg), DMA ceiling with 4 ranks, multi-GPU parity suite once more (kernels changed since the last 2-GPU run).
mkdir -p gpurun_out
O=gpurun_out/r2c11
( time timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 ) > ${O}_bench_4gpu.json 2> ${O}_bench_4gpu.err
tail -c 4500 ${O}_bench_4gpu.json; tail -5 ${O}_bench_4gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29523 scripts/host_dma_ceiling.py > ${O}_dma_4gpu.json 2> ${O}_dma_4gpu.err; cat ${O}_dma_4gpu.json
