#!/bin/bash
# One GPU-box session; everything lands in gpurun_out/ so that a session cut short still leaves what it finished.
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests -m gpu -q --durations=10 > $O/t_all.log 2>&1; echo "all gpu tests rc=$?" | tee -a $O/summary.txt
tail -16 $O/t_all.log
timeout 300 python tools/bench_configs.py 27 > $O/bench_configs.txt 2> $O/bench_configs.err; echo "bench_configs rc=$?" | tee -a $O/summary.txt
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 240 ncu --set full --clock-control none --import-source on -k regex:'k_fm_front_ring' -c 1 -s 2 -o $O/prof_fm_front -f python tools/bench_configs.py 26 > $O/ncu_fm.log 2>&1; echo "ncu fm rc=$?" | tee -a $O/summary.txt
timeout 240 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/launches_bench.log 2>&1; echo "ncu launches rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
