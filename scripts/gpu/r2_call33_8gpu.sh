#!/bin/bash
# Round 2, GPU call 33 (8 GPUs): bench --gpus 8 at HEAD (copy push default, verify, strong block; e2e skipped to save box time).
mkdir -p gpurun_out
O=gpurun_out/r2c33
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29581 bench.py --gpus 8 --steps 20 --warmup 5 --no-e2e ) > ${O}_bench_8gpu.json 2> ${O}_bench_8gpu.err
tail -c 2500 ${O}_bench_8gpu.json; tail -3 ${O}_bench_8gpu.err
