#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-sustained > $O/r2_bench_s17.json 2> $O/r2_bench_s17.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -3 $O/r2_bench_s17.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s17.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    if isinstance(v,dict) and k.startswith('cfg4'): print(k, round(v['ms'],4), round(v['value']), v.get('launches_per_push'), v['kernel'])
print('value', round(d['value']), 'e2e_u8', round(d['e2e_u8']['value']))
PY
