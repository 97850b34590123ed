#!/bin/bash
# Round 2, GPU call 18 (1 GPU): whole GPU suite after the cost-model recalibration (default plans of the small
# test programs may have changed), then the secondary forms (jki, w1d) under the new model.
mkdir -p gpurun_out
O=gpurun_out/r2c18
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-strong"
run() {
  env $2 $B $3 > ${O}_tmp.json 2>> ${O}_err.txt
  python - <<PY
import json
d=json.loads(open("${O}_tmp.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("%-34s %8.4f ms  %.3e upd/s  frac %.3f  %s  verify %s clk %s" % ("$1", d["ms_per_step"], d["value"], r["frac"], r["kernel"], d.get("verify",{}).get("ok"), d["clocks"]["sm_mhz"]))
PY
}
for rep in 1 2; do
run "config2" "A=1" "--config 2"
run "config2-jki model (r3 w12 k32 p5)" "A=1" "--config 2 --variant jki"
run "config2-jki r5 w8 p2" "SFB200_MAX_DEPTH=4 SFB200_ROWS=5 SFB200_WARPS=8 SFB200_PREFETCH=2" "--config 2 --variant jki"
run "config2-jki r4 w8 p5" "SFB200_MAX_DEPTH=4 SFB200_ROWS=4 SFB200_WARPS=8 SFB200_PREFETCH=5" "--config 2 --variant jki"
run "config3-w1d" "A=1" "--config 3 --variant w1d"
done 2>&1 | tee ${O}_variants.txt
