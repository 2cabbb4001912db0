#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 900 python -m pytest tests/test_gpu_tuned_shapes.py -m gpu -q -x > $O/r2_t_s21.log 2>&1; echo "shape tests rc=$?" | tee -a $O/summary.txt
tail -4 $O/r2_t_s21.log
timeout 300 python tools/bench_shapes.py 27 2>&1 | grep -E '"taps": 256' 
