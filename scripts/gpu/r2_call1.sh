#!/bin/bash
# Round 2, GPU call 1 (1 GPU): parity suite with persistent CTAs, then A/B of the persistent schedule
# against the chunk grid on configs 1-3, boundary-path cost, edge-weight sweep.
mkdir -p gpurun_out
O=gpurun_out/r2c1
( time timeout 1500 python -m pytest tests -m gpu -q -x ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e"
for cfg in 1 2 3; do
  $B --config $cfg > ${O}_cfg${cfg}_persistent.json 2> ${O}_cfg${cfg}_persistent.err
  SFB200_PERSISTENT=0 $B --config $cfg > ${O}_cfg${cfg}_chunkgrid.json 2> ${O}_cfg${cfg}_chunkgrid.err
done
SFB200_FASTPATH=0 $B --config 1 > ${O}_cfg1_nofastpath.json 2> ${O}_cfg1_nofastpath.err
for w in 1.05 1.1 1.15 1.2 1.3; do
  SFB200_EDGE_WEIGHT=$w $B --config 1 > ${O}_cfg1_edge$w.json 2> ${O}_cfg1_edge$w.err
done
for w in 1.1 1.2; do
  SFB200_EDGE_WEIGHT=$w $B --config 3 > ${O}_cfg3_edge$w.json 2> ${O}_cfg3_edge$w.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c1_cfg*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s %8.4f ms  %.3e upd/s  frac %.3f  clk %s" % (f.split("r2c1_")[1], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
