#!/bin/bash
# Round 2, GPU call 7 (2 GPUs): multi-GPU parity (in-kernel pushes, copy pushes, one-sided reach), bench --gpus 2
# (weak config 1 + verify against one GPU + configs[4] strong block), DMA ceiling with 2 ranks.
mkdir -p gpurun_out
O=gpurun_out/r2c7
nvidia-smi topo -m > ${O}_topo.txt 2>&1
( time timeout 1500 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x ) > ${O}_pytest.txt 2>&1
tail -15 ${O}_pytest.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 ) > ${O}_bench_2gpu.json 2> ${O}_bench_2gpu.err
tail -c 5000 ${O}_bench_2gpu.json; tail -5 ${O}_bench_2gpu.err
( time SFB200_PEER_PUSH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --no-e2e --no-strong ) > ${O}_bench_2gpu_copypush.json 2> ${O}_bench_2gpu_copypush.err
tail -c 1500 ${O}_bench_2gpu_copypush.json; tail -3 ${O}_bench_2gpu_copypush.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/host_dma_ceiling.py > ${O}_dma_2gpu.json 2> ${O}_dma_2gpu.err; cat ${O}_dma_2gpu.json
timeout 300 python scripts/host_dma_ceiling.py > ${O}_dma_1gpu.json 2> ${O}_dma_1gpu.err; cat ${O}_dma_1gpu.json
