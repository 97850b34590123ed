#!/bin/bash
# 2 GPUs: multi-GPU parity tests, config 4 (2048^3 x 64) strong scaling and config 1 weak scaling at N=2.
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q ) > gpurun_out/c5_pytest_2gpu.txt 2>&1
tail -4 gpurun_out/c5_pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --config 4 --scaling strong --steps 3 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01c_config4_bench_2gpu.json 2> gpurun_out/r01c_config4_bench_2gpu.err
tail -c 700 gpurun_out/r01c_config4_bench_2gpu.json; tail -3 gpurun_out/r01c_config4_bench_2gpu.err
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01c_config1_bench_2gpu.json 2> gpurun_out/r01c_config1_bench_2gpu.err
tail -c 900 gpurun_out/r01c_config1_bench_2gpu.json; tail -3 gpurun_out/r01c_config1_bench_2gpu.err
