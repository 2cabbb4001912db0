#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 300 python -m pytest tests/test_gpu_persistent.py tests/test_gpu_tuned_shapes.py -m gpu -q -x > $O/r2_t_s14.log 2>&1; echo "persist+shapes tests rc=$?" | tee -a $O/summary.txt
tail -5 $O/r2_t_s14.log
timeout 120 python tools/cfg1_probe.py 2>&1 | tail -8
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
tail -c 1500 $O/r2_ncu_launches.log
wc -l $O/r02_launches_bench.csv
