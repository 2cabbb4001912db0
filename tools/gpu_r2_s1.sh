#!/bin/bash
# round-2 GPU session 1 (1 GPU): full GPU test suite, bench with several passes-per-step, host topology probes
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
( nvidia-smi topo -m; echo; lscpu | head -30; echo; numactl -H 2>&1; echo; nproc; cat /proc/meminfo | head -5 ) > $O/r2_topo.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --durations=8 > $O/r2_t_all.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
tail -15 $O/r2_t_all.log
for P in 1 8 64; do
  timeout 300 python bench.py --steps 20 --warmup 5 --passes $P --no-e2e --no-cpu --no-configs > $O/r2_bench_p$P.json 2> $O/r2_bench_p$P.err; echo "bench passes=$P rc=$?" | tee -a $O/summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench.json 2> $O/r2_bench.err; echo "bench full rc=$?" | tee -a $O/summary.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref.json 2> $O/r2_bench_ref.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
for f in $O/r2_bench_p*.json $O/r2_bench.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value',round(d['value']), 'ms/pass', d.get('ms_per_pass'), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks'], 'e2e', (d.get('e2e') or {}).get('value'), 'e2e_u8', (d.get('e2e_u8') or {}).get('value'))
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
