#!/bin/bash
# One GPU-box session; everything lands in gpurun_out/ so that a session cut short still leaves what it finished.
mkdir -p gpurun_out
O=gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "dc_" > $O/t_dc.log 2>&1; echo "dc tests rc=$?" | tee -a $O/summary.txt
tail -3 $O/t_dc.log
timeout 300 python tools/dc_sweep.py 28 27 24 > $O/dc_sweep.txt 2> $O/dc_sweep.err; echo "dc sweep rc=$?" | tee -a $O/summary.txt
cat $O/dc_sweep.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_dc_spec' -c 1 -s 1 -o $O/prof_dc2 -f python tools/dc_probe.py 27 > $O/ncu_dc2.log 2>&1; echo "ncu rc=$?" | tee -a $O/summary.txt
