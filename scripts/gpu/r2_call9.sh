#!/bin/bash
# Round 2, GPU call 9 (1 GPU): program handle through the whole suite, L2 eviction hints, small-grid timing, direct rows.
mkdir -p gpurun_out
O=gpurun_out/r2c9
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
timeout 300 python scripts/small_grid_timing.py > ${O}_small_grid.json 2> ${O}_small_grid.err; cat ${O}_small_grid.json; tail -3 ${O}_small_grid.err
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e --no-strong --no-verify"
for h in 0 1 2 3; do
  SFB200_L2HINT=$h $B --config 1 > ${O}_cfg1_hint$h.json 2> ${O}_cfg1_hint$h.err
  SFB200_L2HINT=$h $B --config 3 > ${O}_cfg3_hint$h.json 2> ${O}_cfg3_hint$h.err
  SFB200_L2HINT=$h $B --config 2 > ${O}_cfg2_hint$h.json 2> ${O}_cfg2_hint$h.err
done
SFB200_MAX_DEPTH=4 SFB200_ROWS=3 SFB200_WARPS=12 SFB200_KS=16 SFB200_PREFETCH=5 $B --config 1 > ${O}_cfg1_explicit.json 2> ${O}_cfg1_explicit.err
SFB200_MAX_DEPTH=4 SFB200_ROWS=3 SFB200_WARPS=12 SFB200_KS=16 SFB200_PREFETCH=5 SFB200_DIRECT=1 $B --config 1 > ${O}_cfg1_direct.json 2> ${O}_cfg1_direct.err
SFB200_MAX_DEPTH=4 SFB200_ROWS=3 SFB200_WARPS=12 SFB200_KS=16 SFB200_PREFETCH=4 $B --config 1 > ${O}_cfg1_p4.json 2> ${O}_cfg1_p4.err
SFB200_MAX_DEPTH=4 SFB200_ROWS=4 SFB200_WARPS=8 SFB200_KS=16 SFB200_PREFETCH=5 $B --config 1 > ${O}_cfg1_r4w8.json 2> ${O}_cfg1_r4w8.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c9_cfg*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-36s %8.4f ms  %.3e upd/s  frac %.3f  clk %s" % (f.split("r2c9_")[1], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
