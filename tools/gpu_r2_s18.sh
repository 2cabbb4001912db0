#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "fm or front or lowrate or chain or state or zero_copy or example or io_edges or pipe or u8" > $O/r2_t_s18.log 2>&1; echo "fm tests rc=$?" | tee -a $O/summary.txt
tail -5 $O/r2_t_s18.log
build/fm_timeline 24 | head -11
build/fm_timeline 27 | head -5
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --no-sustained > $O/r2_bench_s18.json 2> $O/r2_bench_s18.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_s18.json').read().strip().splitlines()[-1])
for k,v in d['configs'].items():
    if isinstance(v,dict) and k.startswith('cfg4'): print(k, round(v['ms'],4), round(v['value']), v.get('launches_per_push'), v['kernel'])
print('value', round(d['value']), 'e2e_u8', round(d['e2e_u8']['value']))
PY
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > $O/r2_san_memcheck3.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; tail -3 $O/r2_san_memcheck3.txt | cut -c1-200
timeout 600 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_targets.py small > $O/r2_san_racecheck3.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; tail -3 $O/r2_san_racecheck3.txt | cut -c1-200
