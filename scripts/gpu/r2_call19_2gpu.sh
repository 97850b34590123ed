#!/bin/bash
# Round 2, GPU call 19 (2 GPUs): multi-GPU parity suite at HEAD (work list, cost model and loop structure changed
# since the last 2-GPU run), bench --gpus 2 with verify and the strong-scaling block.
mkdir -p gpurun_out
O=gpurun_out/r2c19
( time timeout 900 python -m pytest tests/test_distributed_gpu.py -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 ) > ${O}_bench_2gpu.json 2> ${O}_bench_2gpu.err
tail -c 3500 ${O}_bench_2gpu.json; tail -3 ${O}_bench_2gpu.err
