#!/bin/bash
# 8 GPUs: BASELINE configs[4] (Jacobi-3D 2048^3, 64 operators) strong scaling over slabs + NVLink halo pushes.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512"
timeout 500 $TR bench.py --gpus 8 --config 4 --scaling strong --steps 5 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r01c_config4_bench_8gpu.json 2> gpurun_out/r01c_config4_bench_8gpu.err
tail -c 600 gpurun_out/r01c_config4_bench_8gpu.json; tail -3 gpurun_out/r01c_config4_bench_8gpu.err
