#!/bin/bash
# Round 2, GPU call 30 (2 GPUs): which of the new defaults slowed the slab kernels down (4.15 -> 4.60 ms at N = 2).
mkdir -p gpurun_out
O=gpurun_out/r2c30
run() {
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 2 --steps 20 --no-e2e --no-strong --no-verify --no-cpu-baseline > ${O}_tmp.json 2>> ${O}_err.txt
  python - <<PY
import json
d=json.loads(open("${O}_tmp.json").read().strip().splitlines()[-1])
print("%-40s %8.4f ms  %s  clk %s" % ("$1", d["ms_per_step"], d["roofline"]["kernel"], d["clocks"]["sm_mhz"]))
PY
}
run "default" "A=1" 29561
run "bc thread" "SFB200_BC_MODE=thread" 29562
run "reassociate 1" "SFB200_REASSOCIATE=1" 29563
run "sched halving" "SFB200_SCHED=halving" 29564
run "copy push" "SFB200_PEER_PUSH=0" 29565
run "default again" "A=1" 29566
