#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_state.py tests/test_gpu_parity.py -m gpu -q -x -k "state or restore or factor_larger or native_decimator or resample" > $O/r2_t_s3.log 2>&1; echo "tests rc=$?" | tee -a $O/summary.txt
tail -15 $O/r2_t_s3.log
timeout 300 python bench.py --steps 60 --warmup 5 --no-e2e --no-cpu --no-configs > $O/r2_bench_s60.json 2> $O/r2_bench_s60.err; echo "bench s60 rc=$?" | tee -a $O/summary.txt
timeout 300 python bench.py --steps 20 --warmup 5 --no-e2e --no-cpu --no-configs > $O/r2_bench_s20.json 2> $O/r2_bench_s20.err; echo "bench s20 rc=$?" | tee -a $O/summary.txt
for f in $O/r2_bench_s60.json $O/r2_bench_s20.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value',round(d['value']), 'ms/pass', d.get('ms_per_pass'), 'frac', round(d['roofline']['frac'],3), 'copy', d['roofline'].get('copy_GBps_this_box_same_duration'), 'clk', d['clocks'])
    print(' steps', d['ms_by_step_rank0'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
tail -n 3 $O/r2_bench_s60.err
