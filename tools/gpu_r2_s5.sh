#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 1200 python -m pytest tests -m gpu -q -x --durations=8 > $O/r2_t_all2.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
tail -15 $O/r2_t_all2.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs > $O/r2_bench_v3.json 2> $O/r2_bench_v3.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v3.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'sustained',d['sustained']['value'], d['sustained']['roofline_frac'])
PY
timeout 600 python tools/bench_shapes.py 27 > $O/r2_bench_shapes.txt 2> $O/r2_bench_shapes.err; echo "bench_shapes rc=$?" | tee -a $O/summary.txt
cat $O/r2_bench_shapes.txt | cut -c1-220
timeout 1200 compute-sanitizer --tool racecheck --print-limit 40 python tools/sanitize_targets.py small > $O/r2_san_racecheck.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; grep -c "Error: Race" $O/r2_san_racecheck.txt; tail -3 $O/r2_san_racecheck.txt
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > $O/r2_san_memcheck.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; tail -3 $O/r2_san_memcheck.txt
