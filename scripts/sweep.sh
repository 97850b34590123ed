#!/bin/bash
# usage: scripts/sweep.sh <config> "<depths>" "<rows>" [extra env assignments...]
# One bench line (kernel-only) per (depth, rows) combination, summarised on one text line each.
C=$1; DEPTHS=$2; ROWS=$3; shift 3
for d in $DEPTHS; do for r in $ROWS; do
  echo -n "config=$C depth=$d rows=$r $* : "
  env SFB200_MAX_DEPTH=$d SFB200_ROWS=$r "$@" timeout 300 python bench.py --config $C --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
ok=False
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); ok=True
        print('ms/step %.3f  upd/s %.3e  frac %.3f  plan %s clocks %s' % (j['ms_per_step'], j['value'], j['roofline']['frac'], [(p['family'],p['ops']) for p in j['config']['plan']][:4], j['clocks'].get('sm_mhz')))
    elif 'rror' in l: print(l.strip()[:300])
if not ok: print('no result')
"; done; done
