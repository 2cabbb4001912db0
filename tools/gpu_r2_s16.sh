#!/bin/bash
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
timeout 900 python -m pytest tests -m gpu -q -x -k "fm or front or lowrate or chain or state or zero_copy or example or io_edges or pipe" > $O/r2_t_s16.log 2>&1; echo "fm tests rc=$?" | tee -a $O/summary.txt
tail -5 $O/r2_t_s16.log
python tools/chain_probe.py 2>&1 | tail -3
SDR_B200_NO_PDL=1 python tools/chain_probe.py 2>&1 | tail -2
