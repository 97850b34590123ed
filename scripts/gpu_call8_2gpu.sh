#!/bin/bash
# 2 GPUs: multi-GPU parity incl. the exchange-free overlapped host-array call, and the N=2 bench line (e2e).
mkdir -p gpurun_out
( time timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q ) > gpurun_out/c14_pytest_2gpu.txt 2>&1
tail -25 gpurun_out/c14_pytest_2gpu.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r01d_config1_bench_2gpu.json 2> gpurun_out/r01d_config1_bench_2gpu.err
tail -c 900 gpurun_out/r01d_config1_bench_2gpu.json; tail -5 gpurun_out/r01d_config1_bench_2gpu.err
nvidia-smi topo -m > gpurun_out/c14_topo.txt 2>&1
