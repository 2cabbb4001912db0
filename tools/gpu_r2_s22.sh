#!/bin/bash
# final 1-GPU validation of the round: tests, sanitizers, soak, bench (both arms), shape table, ncu launch list + full captures
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 1500 python -m pytest tests -m gpu -q -x --durations=5 > $O/r2_t_final.log 2>&1; echo "gpu tests rc=$?" | tee -a $O/summary.txt
tail -10 $O/r2_t_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/r2_smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt; tail -2 $O/r2_smoke.log
timeout 600 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_targets.py > $O/r2_san_memcheck_final.txt 2>&1; echo "memcheck rc=$?" | tee -a $O/summary.txt; tail -2 $O/r2_san_memcheck_final.txt | cut -c1-160
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_targets.py small > $O/r2_san_synccheck_final.txt 2>&1; echo "synccheck rc=$?" | tee -a $O/summary.txt; tail -2 $O/r2_san_synccheck_final.txt | cut -c1-160
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_targets.py small > $O/r2_san_racecheck_final.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; tail -2 $O/r2_san_racecheck_final.txt | cut -c1-160
timeout 300 python tools/soak.py > $O/r2_soak_final.txt 2>&1; echo "soak rc=$?" | tee -a $O/summary.txt; tail -4 $O/r2_soak_final.txt | cut -c1-200
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_final.json 2> $O/r2_bench_final.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 200 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2_bench_ref_final.json 2>/dev/null; echo "ref rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_final.json').read().strip().splitlines()[-1])
r=json.loads(open('gpurun_out/r2_bench_ref_final.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'sustained',round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'e2e_u8', round(d['e2e_u8']['value']), 'ref', round(r['value']), round(r['e2e_u8']['value']))
for k,v in d['configs'].items():
    print(k, round(v['value']), 'Ms/s', round(v['ms'],4),'ms', 'frac', round(v['roofline']['frac'],3), 'fp32', round(v['roofline'].get('fp32_frac',0),3), v['kernel'], v.get('launches_per_push'))
print({k:(round(v['value']) if isinstance(v,dict) else v) for k,v in d['pipes_mode'].items() if k!='note'})
PY
timeout 400 python tools/bench_shapes.py 27 > $O/r02_bench_shapes.txt 2>&1; echo "shapes rc=$?" | tee -a $O/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_dec_ring|k_fir_ring|k_res_ring|k_fm_front_ring|k_fm_lowrate|k_dc_spec_tiles' -c 15 -o $O/r02_prof python tools/ncu_targets.py 27 > $O/r2_ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $O/summary.txt; tail -2 $O/r2_ncu_full.log
ls -la $O/*.ncu-rep
