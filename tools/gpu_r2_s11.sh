#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=6 > $O/r2_t_all4.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
tail -12 $O/r2_t_all4.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_v6.json 2> $O/r2_bench_v6.err; echo "bench rc=$?" | tee -a $O/summary.txt
tail -3 $O/r2_bench_v6.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v6.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'sustained',d['sustained']['value'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'])
for k,v in d['configs'].items():
    print(k, round(v['value']), 'Ms/s', round(v['ms'],4),'ms', 'frac', round(v['roofline']['frac'],3), 'fp32', round(v['roofline'].get('fp32_frac',0),3), v['kernel'], v.get('launches_per_push'))
for k,v in d['pipes_mode'].items():
    if isinstance(v, dict): print(k, round(v['value']), v.get('ms'))
PY
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref2.json 2>/dev/null; echo "ref rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke2.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > $O/r2_san_memcheck2.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; tail -4 $O/r2_san_memcheck2.txt | cut -c1-300
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python tools/sanitize_targets.py small > $O/r2_san_racecheck2.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; grep "Error: Race" $O/r2_san_racecheck2.txt | sed 's/+0x[0-9a-f]*//' | cut -c1-200 | sort | uniq -c; tail -2 $O/r2_san_racecheck2.txt | cut -c1-300
timeout 600 python tools/soak.py 1000 26 > $O/r2_soak2.txt 2>&1; echo "soak rc=$?" | tee -a $O/summary.txt; tail -2 $O/r2_soak2.txt
