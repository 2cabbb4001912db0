#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x > $O/r2_t_s24.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
tail -6 $O/r2_t_s24.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-sustained > $O/r2_bench_s24.json 2> $O/r2_bench_s24.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s24.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    if isinstance(v,dict) and k.startswith('cfg4'): print(k, round(v['ms'],4), round(v['value']), v.get('launches_per_push'), v['kernel'])
print('value', round(d['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']))
print({k:(round(v['value']) if isinstance(v,dict) else v) for k,v in d['pipes_mode'].items() if k!='note'})
PY
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > $O/r2_san_memcheck_final.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; tail -1 $O/r2_san_memcheck_final.txt
timeout 200 python tools/persist_probe2.py vec 28 60 | tail -1
