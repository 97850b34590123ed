#!/bin/bash
# Round 2, GPU call 21 (1 GPU): small-grid planning after the latency-model fix (size sweep, smoke under ncu),
# whole GPU suite (plans of the small test programs changed).
mkdir -p gpurun_out
O=gpurun_out/r2c21
timeout 600 python scripts/small_grid_sizes.py 2>&1 | tee ${O}_small_grid_sizes.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_smoke_launches.csv python __graft_entry__.py smoke > ${O}_smoke.txt 2>&1
tail -2 ${O}_smoke.txt
python - <<'PY'
import csv, io
lines = open("gpurun_out/r2c21_smoke_launches.csv").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    print(r["Kernel Name"][:40], r["Metric Value"], r["Metric Unit"])
PY
( time timeout 1500 python -m pytest tests -m gpu -q ) > ${O}_pytest.txt 2>&1
tail -4 ${O}_pytest.txt
