#!/bin/bash
# Round 2, GPU call 3 (1 GPU): persistent CTAs with a dynamic work queue vs the chunk grid; ncu captures of both.
mkdir -p gpurun_out
O=gpurun_out/r2c3
( time timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q -x -k "persistent or chunk_grid or planned or pipelined" ) > ${O}_pytest.txt 2>&1
tail -5 ${O}_pytest.txt
B="timeout 300 python bench.py --steps 20 --no-cpu-baseline --no-e2e"
for cfg in 1 2 3; do
  $B --config $cfg > ${O}_cfg${cfg}_persistent.json 2> ${O}_cfg${cfg}_persistent.err
  SFB200_PERSISTENT=0 $B --config $cfg > ${O}_cfg${cfg}_chunkgrid.json 2> ${O}_cfg${cfg}_chunkgrid.err
done
python - <<'PY'
import glob, json
for f in sorted(glob.glob("gpurun_out/r2c3_cfg*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s %8.4f ms  %.3e upd/s  frac %.3f  clk %s" % (f.split("r2c3_")[1], d["ms_per_step"], d["value"], d["roofline"]["frac"], d["clocks"]["sm_mhz"]))
    except Exception as e:
        print(f, "FAILED", e)
PY
for cfg in 1 2; do
  for mode in 1 0; do
    SFB200_PERSISTENT=$mode timeout 600 ncu --set full --clock-control none --import-source on -k regex:sf_ -s 3 -c 1 \
      -o ${O}_cfg${cfg}_p${mode}_full -f python bench.py --config $cfg --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > /dev/null 2>&1
    python scripts/ncu_summary.py ${O}_cfg${cfg}_p${mode}_full.ncu-rep > ${O}_cfg${cfg}_p${mode}_summary.txt
    grep -E "gpu__time_duration|DRAM traffic|lts__t_bytes|warps_active|issue_active|stall" ${O}_cfg${cfg}_p${mode}_summary.txt
  done
done
