#!/bin/bash
# Round 2, GPU call 40 (1 GPU): shorter chunks for 2-D rows as the default: whole GPU suite, config 3 bench + ncu set.
mkdir -p gpurun_out
O=gpurun_out/r2c40
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
TAG=r02f; OUT=gpurun_out; c=3
timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-strong > $OUT/${TAG}_config${c}_bench.json 2> $OUT/${TAG}_config${c}_bench.err
tail -c 700 $OUT/${TAG}_config${c}_bench.json; echo
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/${TAG}_config${c}_launches.csv python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-verify > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sf_ -s 3 -c 1 -o $OUT/${TAG}_config${c}_full -f python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-verify > /dev/null 2>&1
ALG=$(python -c "import json;print(json.load(open('$OUT/${TAG}_config${c}_bench.json'))['roofline']['algorithmic_bytes_per_launch'])")
python scripts/ncu_summary.py $OUT/${TAG}_config${c}_full.ncu-rep $ALG > $OUT/${TAG}_config${c}_ncu_full_summary.txt
python scripts/ncu_source.py $OUT/${TAG}_config${c}_full.ncu-rep --top 16 >> $OUT/${TAG}_config${c}_ncu_full_summary.txt
grep -E "gpu__time_duration|DRAM traffic|traffic /|stall" $OUT/${TAG}_config${c}_ncu_full_summary.txt
timeout 300 python bench.py --config 3 --variant w1d --steps 20 --no-cpu-baseline --no-e2e --no-strong > ${O}_w1d.json 2>/dev/null; python -c "
import json; d=json.loads(open('${O}_w1d.json').read().strip().splitlines()[-1]); print('w1d', d['ms_per_step'], d['roofline']['frac'], d.get('verify',{}).get('ok'))"
