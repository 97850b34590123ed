#!/bin/bash
# Round 2, GPU call 8 (2 GPUs): in-kernel halo push moved behind the segment (no extra registers in the streamed loop).
mkdir -p gpurun_out
O=gpurun_out/r2c8
( time timeout 1500 python -m pytest tests/test_distributed_gpu.py -m gpu -q -x ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 ) > ${O}_bench_2gpu.json 2> ${O}_bench_2gpu.err
tail -c 5000 ${O}_bench_2gpu.json; tail -5 ${O}_bench_2gpu.err
( time SFB200_PEER_PUSH=0 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --no-e2e ) > ${O}_bench_2gpu_copypush.json 2> ${O}_bench_2gpu_copypush.err
tail -c 2500 ${O}_bench_2gpu_copypush.json; tail -3 ${O}_bench_2gpu_copypush.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --no-e2e --no-cpu-baseline ) > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err
tail -c 2500 ${O}_bench_1gpu.json
