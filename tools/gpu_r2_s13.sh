#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > $O/r2_t_final.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
tail -12 $O/r2_t_final.log
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > $O/r2_san_memcheck2.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; tail -3 $O/r2_san_memcheck2.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref_final.json 2>/dev/null; echo "ref rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_ref_final.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'sustained',round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), 'ref', round(r['value']), round(r['e2e_u8']['value']))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
