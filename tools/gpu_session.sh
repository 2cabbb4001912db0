#!/bin/bash
# One GPU-box validation session (run as `gpurun -- 'bash tools/gpu_session.sh'`): the whole GPU test suite, the headline
# bench, the per-config / per-stage measurements and smoke().  Everything lands in gpurun_out/ so that a session cut
# short still leaves what it finished.
mkdir -p gpurun_out
O=gpurun_out
rm -f $O/summary.txt
timeout 600 python -m pytest tests -m gpu -q --durations=5 > $O/t_all.log 2>&1; echo "all gpu tests rc=$?" | tee -a $O/summary.txt
tail -9 $O/t_all.log
timeout 300 python bench.py --steps 20 --warmup 3 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?" | tee -a $O/summary.txt
timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref.json 2> $O/bench_ref.err; echo "bench ref rc=$?" | tee -a $O/summary.txt
timeout 300 python tools/bench_configs.py 27 > $O/bench_configs.txt 2> $O/bench_configs.err; echo "bench_configs rc=$?" | tee -a $O/summary.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke rc=$?" | tee -a $O/summary.txt
