for d in 1 2 3 4; do for r in 2 3 4; do
  echo "depth=$d rows=$r"; SFB200_MAX_DEPTH=$d SFB200_ROWS=$r timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-e2e 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        j=json.loads(l); print('  ms/step %.3f  upd/s %.3e  frac %.3f  plan %s'%(j['ms_per_step'], j['value'], j['roofline']['frac'], [(p['family'],p['ops']) for p in j['config']['plan']][:3]))
    elif 'Error' in l or 'error' in l: print(l.strip()[:200])
"; done; done
