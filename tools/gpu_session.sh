#!/bin/bash
# One GPU-box session; everything lands in gpurun_out/ so that a session cut short still leaves what it finished.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/res_sweep.txt $O/summary.txt
timeout 600 python -m pytest tests -m gpu -q > $O/t_all.log 2>&1; echo "all gpu tests rc=$?" | tee -a $O/summary.txt
tail -3 $O/t_all.log
for s in 2 1; do SDR_B200_RES_S=$s timeout 100 python tools/res_sweep.py >> $O/res_sweep.txt 2>> $O/res_sweep.err; done
cat $O/res_sweep.txt
SDR_B200_RES_S=1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "resampler" 2>&1 | tail -2
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
