#!/bin/bash
# One B200: for BASELINE configs 1-3 the bench line, the ncu launch list of the same command and one
# `ncu --set full` capture of the dominant kernel; condensed summaries and the DRAM traffic per launch
# (profiles/traffic.json, read by bench.py) are derived from the captures.  Outputs: gpurun_out/<tag>_*.
TAG=${1:-r01}
OUT=gpurun_out
mkdir -p $OUT
for c in 1 2 3; do
  timeout 900 python bench.py --config $c --steps 20 --warmup 3 --no-strong > $OUT/${TAG}_config${c}_bench.json 2> $OUT/${TAG}_config${c}_bench.err
  tail -c 600 $OUT/${TAG}_config${c}_bench.json; echo
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
      --log-file $OUT/${TAG}_config${c}_launches.csv \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-verify > /dev/null 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:sf_ -s 3 -c 1 \
      -o $OUT/${TAG}_config${c}_full -f \
      python bench.py --config $c --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-strong --no-verify > /dev/null 2>&1
  ALG=$(python -c "import json;print(json.load(open('$OUT/${TAG}_config${c}_bench.json'))['roofline']['algorithmic_bytes_per_launch'])")
  python scripts/ncu_summary.py $OUT/${TAG}_config${c}_full.ncu-rep $ALG > $OUT/${TAG}_config${c}_ncu_full_summary.txt
  python scripts/ncu_source.py $OUT/${TAG}_config${c}_full.ncu-rep --top 16 >> $OUT/${TAG}_config${c}_ncu_full_summary.txt
  grep -E "gpu__time_duration|DRAM traffic|traffic /|stall" $OUT/${TAG}_config${c}_ncu_full_summary.txt
done
python - <<PY
import json, re
out = {}
for c in (1, 2, 3):
    try:
        text = open("$OUT/${TAG}_config%d_ncu_full_summary.txt" % c).read()
        m = re.search(r"DRAM traffic per launch: ([0-9.e+]+) B", text)
        k = re.search(r"kernel: (\S+)", text)
        out[str(c)] = {"dram_bytes_per_launch": float(m.group(1)), "kernel": k.group(1), "source": "ncu --set full, ${TAG}"}
    except Exception as exc:
        print("config", c, exc)
json.dump(out, open("$OUT/${TAG}_traffic.json", "w"), indent=1)
print(out)
PY
ls -la $OUT | tail -25
