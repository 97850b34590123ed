#!/bin/bash
# Round 2, GPU call 31 (4 GPUs): in-kernel push vs copy push of the halo planes at N = 4 under the HEAD defaults.
mkdir -p gpurun_out
O=gpurun_out/r2c31
run() {
  env $2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port $3 bench.py --gpus 4 --steps 20 --no-e2e --no-verify --no-cpu-baseline $4 > ${O}_tmp.json 2>> ${O}_err.txt
  python - <<PY
import json
d=json.loads(open("${O}_tmp.json").read().strip().splitlines()[-1]); s=d.get("strong") or {}
print("%-28s %8.4f ms  %s  clk %s   strong %s ms eff %s" % ("$1", d["ms_per_step"], d["roofline"]["kernel"], d["clocks"]["sm_mhz"], s.get("ms_per_step"), s.get("parallel_efficiency")))
PY
}
run "copy push" "SFB200_PEER_PUSH=0" 29571 ""
run "in-kernel push" "A=1" 29572 ""
run "copy push again" "SFB200_PEER_PUSH=0" 29573 "--no-strong"
run "in-kernel, reassociate 1" "SFB200_REASSOCIATE=1" 29574 "--no-strong"
