#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs > $O/r2_bench_v2.json 2> $O/r2_bench_v2.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v2.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'clk',d['clocks'])
print('steps',d['ms_by_step_rank0'])
print('sustained',d['sustained'])
PY
timeout 600 python tools/soak.py 1000 26 > $O/r2_soak.txt 2>&1; echo "soak rc=$?" | tee -a $O/summary.txt; tail -4 $O/r2_soak.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > $O/r2_san_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; tail -6 $O/r2_san_memcheck.txt
timeout 1200 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_targets.py small > $O/r2_san_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; tail -6 $O/r2_san_racecheck.txt
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_targets.py small > $O/r2_san_synccheck.txt 2>&1; echo "synccheck rc=$?" | tee -a $O/summary.txt; tail -4 $O/r2_san_synccheck.txt
