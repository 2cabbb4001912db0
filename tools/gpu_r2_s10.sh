#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 400 python -m pytest tests/test_gpu_persistent.py -m gpu -q -x > $O/r2_t_persist.log 2>&1; echo "persist tests rc=$?" | tee -a $O/summary.txt
tail -30 $O/r2_t_persist.log
timeout 300 python - <<'PY' 2>&1 | tail -20
import sys, json, ctypes as C, time
sys.path.insert(0, '.')
import bench, sdr_b200
from sdr_b200 import _lib as L
class A: log2n=26
ctx = sdr_b200.default_context()
dec = sdr_b200.cudaDecimatorC(8, bench.design_taps(), ctx=ctx, sizeMultiple=4)
print(json.dumps(bench.run_pipes_mode(A, ctx, dec, sdr_b200, L), indent=1))
PY
