#!/bin/bash
# Round 2, GPU call 4 (1 GPU): persistent CTAs with the factoring work list vs the chunk grid; boundary-code variants.
mkdir -p gpurun_out
O=gpurun_out/r2c4
( time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "persistent or chunk_grid or planned or pipelined" ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e"
for cfg in 1 2 3; do
  $B --config $cfg > ${O}_cfg${cfg}_persistent.json 2> ${O}_cfg${cfg}_persistent.err
  SFB200_PERSISTENT=0 $B --config $cfg > ${O}_cfg${cfg}_chunkgrid.json 2> ${O}_cfg${cfg}_chunkgrid.err
done
SFB200_BC_MODE=cta $B --config 1 > ${O}_cfg1_persistent_bccta.json 2> ${O}_cfg1_persistent_bccta.err
SFB200_BC_MODE=cta SFB200_PERSISTENT=0 $B --config 1 > ${O}_cfg1_chunkgrid_bccta.json 2> ${O}_cfg1_chunkgrid_bccta.err
$B --config 1 > ${O}_cfg1_persistent_again.json 2> ${O}_cfg1_persistent_again.err
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c4_cfg*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s %8.4f ms  %.3e upd/s  frac %.3f  clk %s" % (f.split("r2c4_")[1], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
SFB200_PERSISTENT=1 timeout 600 ncu --set full --clock-control none --import-source on -k regex:sf_ -s 3 -c 1 \
      -o ${O}_cfg1_p1_full -f python bench.py --config 1 --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
python scripts/ncu_summary.py ${O}_cfg1_p1_full.ncu-rep > ${O}_cfg1_p1_summary.txt
grep -E "gpu__time_duration|DRAM traffic|lts__t_bytes|warps_active|issue_active|stall" ${O}_cfg1_p1_summary.txt
