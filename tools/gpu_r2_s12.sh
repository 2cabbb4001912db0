#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 600 python -m pytest tests/test_gpu_persistent.py tests/test_gpu_tuned_shapes.py -m gpu -q -x > $O/r2_t_s12.log 2>&1; echo "tests rc=$?" | tee -a $O/summary.txt; tail -4 $O/r2_t_s12.log
timeout 600 python tools/bench_shapes.py 27 > $O/r2_bench_shapes2.txt 2> $O/r2_bench_shapes2.err; echo "bench_shapes rc=$?" | tee -a $O/summary.txt
grep '"D": 4' $O/r2_bench_shapes2.txt | cut -c1-220
timeout 900 compute-sanitizer --tool racecheck --print-limit 40 python tools/sanitize_targets.py small > $O/r2_san_racecheck2.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/summary.txt; grep "Error: Race" $O/r2_san_racecheck2.txt | sed 's/+0x[0-9a-f]*//' | cut -c1-160 | sort | uniq -c; tail -2 $O/r2_san_racecheck2.txt | cut -c1-300
timeout 600 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_targets.py small > $O/r2_san_synccheck2.txt 2>&1; echo "synccheck rc=$?" | tee -a $O/summary.txt; tail -2 $O/r2_san_synccheck2.txt | cut -c1-200
