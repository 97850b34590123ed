#!/bin/bash
# Round 2, GPU call 2 (1 GPU): wave-structured persistent schedule vs chunk grid, configs 1-3.
mkdir -p gpurun_out
O=gpurun_out/r2c2
( time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "persistent or chunk_grid or planned or pipelined" ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e"
for cfg in 1 2 3; do
  $B --config $cfg > ${O}_cfg${cfg}_persistent.json 2> ${O}_cfg${cfg}_persistent.err
  SFB200_PERSISTENT=0 $B --config $cfg > ${O}_cfg${cfg}_chunkgrid.json 2> ${O}_cfg${cfg}_chunkgrid.err
done
$B --config 1 --steps 100 > ${O}_cfg1_persistent_100.json 2> ${O}_cfg1_persistent_100.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c2_cfg*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s %8.4f ms  %.3e upd/s  frac %.3f  clk %s" % (f.split("r2c2_")[1], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
