#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=8 > $O/r2_t_all3.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
tail -25 $O/r2_t_all3.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_v4.json 2> $O/r2_bench_v4.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -3 $O/r2_bench_v4.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v4.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'sustained',d['sustained']['value'], d['sustained']['roofline_frac'])
print('e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'])
for k,v in d['configs'].items():
    print(k, round(v['value']), 'Ms/s', round(v['ms'],4),'ms', 'frac', round(v['roofline']['frac'],3), 'fp32', round(v['roofline'].get('fp32_frac',0),3), v['kernel'], v.get('launches_per_push'))
print(json.dumps(d['pipes_mode'], indent=1))
PY
