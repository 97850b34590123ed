#!/bin/bash
# Round 2, GPU call 16 (1 GPU): end-to-end host-array call -- ramped vs equal pieces, 16 vs 32 pieces.
mkdir -p gpurun_out
O=gpurun_out/r2c16
B="timeout 300 python bench.py --steps 5 --no-cpu-baseline --no-verify --no-strong"
for c in 1 3 2; do
for v in "0 16" "1 16" "1 32" "0 32" "1 16"; do
  set -- $v
  SFB200_PIPELINE_RAMP=$1 SFB200_PIPELINE_PIECES=$2 $B --config $c > ${O}_tmp.json 2>> ${O}_err.txt
  python - <<PY
import json
d=json.loads(open("${O}_tmp.json").read().strip().splitlines()[-1]); e=d["e2e"]
print("config $c ramp $1 pieces $2: e2e %.2f ms (numpy arrays) %.2f ms (pinned)  pcie frac %s" % (e["ms_per_step"], e.get("pinned",{}).get("ms_per_step", float("nan")), e.get("roofline",{}).get("frac")))
PY
done
done 2>&1 | tee ${O}_e2e.txt
