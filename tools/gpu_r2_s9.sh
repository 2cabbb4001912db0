#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
# PDL on/off at the per-GPU chunk size of an 8-GPU run (2^25 samples, 24 passes per step) and at the headline size
for L2N in 25 28; do for V in pdl nopdl; do
  if [ $V = nopdl ]; then export SDR_B200_NO_PDL=1; else unset SDR_B200_NO_PDL; fi
  P=3; [ $L2N = 25 ] && P=24
  timeout 200 python bench.py --steps 20 --warmup 5 --log2n $L2N --passes $P --no-e2e --no-cpu --no-configs --no-sustained > $O/r2_pdl_${L2N}_$V.json 2>/dev/null
  python - "$O/r2_pdl_${L2N}_$V.json" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(sys.argv[1], 'ms/pass', round(d['ms_per_pass'],5), 'value', round(d['value']))
PY
  sleep 1
done; done
unset SDR_B200_NO_PDL
timeout 300 python tools/ab_kernels.py build/ab/libsdr_b200_r1.so sdr_b200/lib/libsdr_b200.so > $O/r2_ab2.txt 2>&1; tail -3 $O/r2_ab2.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2_bench_v5.json 2> $O/r2_bench_v5.err; echo "bench rc=$?" | tee -a $O/summary.txt
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2_bench_v5.json').read().strip().splitlines()[-1])
print('value',round(d['value']),'frac',round(d['roofline']['frac'],3),'sustained',d['sustained']['value'], 'e2e', d['e2e']['value'], 'e2e_u8', d['e2e_u8']['value'])
for k,v in d['configs'].items():
    print(k, round(v['value']), 'Ms/s', round(v['ms'],4),'ms', 'frac', round(v['roofline']['frac'],3), 'fp32', round(v['roofline'].get('fp32_frac',0),3), v['kernel'], v.get('launches_per_push'))
PY
# launch list of the bench (every launch with its device time)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
# full-set captures of one launch of every tuned kernel
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:'k_dec_ring|k_fir_ring|k_res_ring|k_fm_front_ring|k_fm_lowrate|k_dc_spec_tiles' -c 14 -o $O/r02_prof python tools/ncu_targets.py 27 > $O/r2_ncu_full.log 2>&1; echo "ncu full rc=$?" | tee -a $O/summary.txt; tail -3 $O/r2_ncu_full.log
ls -la $O/*.ncu-rep
