#!/bin/bash
# One GPU-box session; everything lands in gpurun_out/ so that a session cut short still leaves what it finished.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/dc_modes.txt
timeout 900 python -m pytest tests -m gpu -q --durations=5 > $O/t_all.log 2>&1; echo "all gpu tests rc=$?" | tee -a $O/summary.txt
tail -9 $O/t_all.log
for mode in 2 3; do
SDR_B200_DC_MODE=$mode DC_SWEEP_SHORT=1 timeout 200 python tools/dc_sweep.py 28 27 >> $O/dc_modes.txt 2>> $O/dc_modes.err; echo "mode $mode rc=$?" | tee -a $O/summary.txt
done
cat $O/dc_modes.txt
timeout 300 python tools/dc_sweep.py 28 > $O/dc_sweep.txt 2> $O/dc_sweep.err; echo "dc sweep rc=$?" | tee -a $O/summary.txt
timeout 300 python tools/bench_configs.py 27 > $O/bench_configs.txt 2> $O/bench_configs.err; echo "bench_configs rc=$?" | tee -a $O/summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_dc_spec' -c 1 -s 1 -o $O/prof_dc4 -f python tools/dc_probe.py 28 > $O/ncu_dc4.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
