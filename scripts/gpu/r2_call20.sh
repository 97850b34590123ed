#!/bin/bash
# Round 2, GPU call 20 (1 GPU): smoke() under the ncu launch list (kernel durations of the 32^3 case, planned vs
# one-operator kernels), then the exact default bench command the driver runs (timed), and the reference arm.
mkdir -p gpurun_out
O=gpurun_out/r2c20
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file ${O}_smoke_launches.csv python __graft_entry__.py smoke > ${O}_smoke.txt 2>&1
tail -3 ${O}_smoke.txt
python - <<'PY'
import csv, io
lines = open("gpurun_out/r2c20_smoke_launches.csv").read().splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
for r in csv.DictReader(io.StringIO("\n".join(lines[start:]))):
    print(r["Kernel Name"][:40], r["Grid Size"] if "Grid Size" in r else "", r["Metric Value"], r["Metric Unit"])
PY
( time timeout 1200 python bench.py ) > ${O}_bench_default.json 2> ${O}_bench_default.err
tail -c 3000 ${O}_bench_default.json; tail -4 ${O}_bench_default.err
( time timeout 1200 python bench.py --impl reference ) > ${O}_bench_reference.json 2> ${O}_bench_reference.err
tail -c 1200 ${O}_bench_reference.json; tail -4 ${O}_bench_reference.err
