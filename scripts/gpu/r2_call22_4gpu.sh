#!/bin/bash
# Round 2, GPU call 22 (4 GPUs): bench --gpus 4 at HEAD (longest-first work list, recalibrated planner): weak config 1
# with verify against one GPU, the configs[4] strong block, and the one-GPU weak baseline of the same box.
mkdir -p gpurun_out
O=gpurun_out/r2c22
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 4 --steps 20 --warmup 5 ) > ${O}_bench_4gpu.json 2> ${O}_bench_4gpu.err
tail -c 2600 ${O}_bench_4gpu.json; tail -3 ${O}_bench_4gpu.err
( time timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-e2e --no-cpu-baseline --no-strong ) > ${O}_bench_1gpu.json 2> ${O}_bench_1gpu.err
tail -c 900 ${O}_bench_1gpu.json
