#!/bin/bash
# Round 2, GPU call 17 (1 GPU): what the cost model alone (SFB200_TUNED=0) delivers next to the measured plans.
mkdir -p gpurun_out
O=gpurun_out/r2c17
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-strong --no-verify"
run() {  # label, env, args
  env $2 $B $3 > ${O}_tmp.json 2>> ${O}_err.txt
  python - <<PY
import json
d=json.loads(open("${O}_tmp.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("%-28s %8.4f ms  %.3e upd/s  frac %.3f  %s  plan %s  clk %s" % ("$1", d["ms_per_step"], d["value"], r["frac"], r["kernel"], d["plan"][:3], d["clocks"]["sm_mhz"]))
PY
}
for rep in 1 2; do
run "config1 tuned" "A=1" "--config 1"
run "config1 model" "SFB200_TUNED=0" "--config 1"
run "config2 tuned" "A=1" "--config 2"
run "config2 model" "SFB200_TUNED=0" "--config 2"
run "config2-jki default" "A=1" "--config 2 --variant jki"
run "config3 tuned" "A=1" "--config 3"
run "config3 model" "SFB200_TUNED=0" "--config 3"
run "config3-w1d default" "A=1" "--config 3 --variant w1d"
done 2>&1 | tee ${O}_model_vs_tuned.txt
