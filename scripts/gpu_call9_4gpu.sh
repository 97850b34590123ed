#!/bin/bash
# 4 GPUs: weak-scaling bench line of config 1 incl. the exchange-free overlapped host-array call (e2e).
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29514"
timeout 400 $TR bench.py --gpus 4 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_config1_bench_4gpu.json 2> gpurun_out/r01d_config1_bench_4gpu.err
tail -c 700 gpurun_out/r01d_config1_bench_4gpu.json; tail -3 gpurun_out/r01d_config1_bench_4gpu.err
