#!/bin/bash
# Round 2, GPU call 41 (2 GPUs): multi-GPU parity suite at HEAD (one-warp 2-D CTAs, short chunks) + config 3 on 2 GPUs.
mkdir -p gpurun_out
O=gpurun_out/r2c41
( time timeout 600 python -m pytest tests/test_distributed_gpu.py -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -3 ${O}_pytest.txt
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29591 bench.py --gpus 2 --config 3 --steps 10 --no-e2e --no-cpu-baseline > ${O}_cfg3_2gpu.json 2> ${O}_cfg3_2gpu.err
python -c "
import json; d=json.loads(open('${O}_cfg3_2gpu.json').read().strip().splitlines()[-1]); print('config 3 on 2 GPUs', d['ms_per_step'], d['value'], d['roofline']['frac'], d.get('verify'))"; tail -2 ${O}_cfg3_2gpu.err
