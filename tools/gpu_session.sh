#!/bin/bash
# One GPU-box session; everything lands in gpurun_out/ so that a session cut short still leaves what it finished.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "real_filter" > $O/t_ffa.log 2>&1; echo "ffa tests rc=$?" | tee -a $O/summary.txt
tail -12 $O/t_ffa.log
timeout 300 python tools/bench_configs.py 27 > $O/bench_configs.txt 2> $O/bench_configs.err; echo "bench_configs rc=$?" | tee -a $O/summary.txt
head -3 $O/bench_configs.txt | cut -c1-300
SDR_B200_FIR_FFA=1 timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_fir_r_ffa_ring' -c 1 -s 2 -o $O/prof_ffa -f python tools/bench_configs.py 26 > $O/ncu_ffa.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
