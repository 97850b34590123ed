#!/bin/bash
# Round 2, GPU call 14 (1 GPU): longest-first work list, unroll 12, halo-warp skip, split loops (config 1);
# split barrier / split loops / deeper TMA ring (config 3); hdiff check after the boundary fix-up change.
mkdir -p gpurun_out
O=gpurun_out/r2c14
K=d4r3w12p5
timeout 900 python scripts/sweep_variants.py --config 1 --steps 10 --repeat 3 $K SFB200_SCHED=lpt:$K SFB200_UNROLL=12,SFB200_UNROLL_MULT=2:$K SFB200_SCHED=lpt,SFB200_UNROLL=12,SFB200_UNROLL_MULT=2:$K SFB200_HALO_SKIP=1:$K SFB200_SPLITLOOP=1:$K SFB200_UNROLL=12,SFB200_UNROLL_MULT=2,SFB200_ST64=1:$K > ${O}_sweep1.txt 2>&1
grep -A20 medians ${O}_sweep1.txt; grep -i "differ\|fail" ${O}_sweep1.txt | head
K3=d8v4w2p5
timeout 900 python scripts/sweep_variants.py --config 3 --steps 5 --repeat 3 $K3 SFB200_SPLITBAR=1:$K3 SFB200_SPLITLOOP=1:$K3 SFB200_SPLITBAR=1,SFB200_SPLITLOOP=1:$K3 SFB200_UNROLL=12:d8v4w2p11 SFB200_SPLITBAR=1,SFB200_UNROLL=12:d8v4w2p11 d8v4w2p2 > ${O}_sweep3.txt 2>&1
grep -A20 medians ${O}_sweep3.txt; grep -i "differ\|fail" ${O}_sweep3.txt | head
timeout 300 python bench.py --config 2 --steps 50 --no-cpu-baseline --no-e2e > ${O}_cfg2.json 2> ${O}_cfg2.err
python -c "
import json; d=json.loads(open('${O}_cfg2.json').read().strip().splitlines()[-1]); print('config2', d['ms_per_step'], d['roofline']['frac'], d['roofline']['kernel'], d.get('verify',{}).get('ok'))"
( SFB200_HALO_SKIP=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "plan_variants_3d" ) > ${O}_pytest_haloskip.txt 2>&1; tail -3 ${O}_pytest_haloskip.txt
( SFB200_SPLITLOOP=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "plan_variants" ) > ${O}_pytest_splitloop.txt 2>&1; tail -3 ${O}_pytest_splitloop.txt
( SFB200_SPLITBAR=1 timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q -k "plan_variants or pair_sync_bit" ) > ${O}_pytest_splitbar.txt 2>&1; tail -3 ${O}_pytest_splitbar.txt
