#!/bin/bash
# Round 2, GPU call 10 (1 GPU): re-associated odd-k taps A/B, copy boundaries in fused passes, then the round's
# profile set (bench line, launch list, ncu --set full) for configs 1-3.
mkdir -p gpurun_out
O=gpurun_out/r2c10
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-strong --no-verify"
for r in 1 0 1 0; do
  SFB200_REASSOCIATE=$r $B --config 1 >> ${O}_cfg1_reassoc$r.json 2>> ${O}_cfg1_reassoc$r.err
done
SFB200_REASSOCIATE=2 $B --config 1 > ${O}_cfg1_reassoc2.json 2> ${O}_cfg1_reassoc2.err
for r in 1 0 1 0; do
  SFB200_REASSOCIATE=$r $B --config 3 >> ${O}_cfg3_reassoc$r.json 2>> ${O}_cfg3_reassoc$r.err
done
SFB200_REASSOCIATE=0 $B --config 2 > ${O}_cfg2_reassoc0.json 2> ${O}_cfg2_reassoc0.err
SFB200_REASSOCIATE=1 $B --config 2 > ${O}_cfg2_reassoc1.json 2> ${O}_cfg2_reassoc1.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c10_cfg*.json")):
    for line in open(f).read().strip().splitlines():
        try:
            d = json.loads(line)
            print("%-36s %8.4f ms  %.3e upd/s  frac %.3f  clk %s" % (f.split("r2c10_")[1], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
        except Exception as e:
            print(f, "FAILED", e)
PY
bash scripts/profile_round.sh r02 > ${O}_profile_round.txt 2>&1
tail -30 ${O}_profile_round.txt
