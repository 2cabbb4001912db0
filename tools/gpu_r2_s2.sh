#!/bin/bash
# round-2 GPU session 2 (N GPUs, default 2): multi-rank parity tests + scaling bench N=1..NG
NG=${1:-2}
mkdir -p gpurun_out; O=gpurun_out
rm -f $O/summary.txt
nvidia-smi topo -m > $O/r2_topo_${NG}gpu.txt 2>&1
timeout 900 python -m pytest tests/test_multi_rank.py tests/test_gpu_parity.py -m gpu -q -x -k "sharded or resample or ring_kernel" > $O/r2_t_mg_${NG}.log 2>&1; echo "mg tests rc=$?" | tee -a $O/summary.txt
tail -5 $O/r2_t_mg_${NG}.log
N=1
while [ $N -le $NG ]; do
  if [ $N -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-configs --no-cpu > $O/r2_scale_n$N.json 2> $O/r2_scale_n$N.err
  else
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 20 --warmup 5 > $O/r2_scale_n$N.json 2> $O/r2_scale_n$N.err
  fi
  echo "bench N=$N rc=$?" | tee -a $O/summary.txt
  N=$((N*2))
done
for f in $O/r2_scale_n*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], 'value',round(d['value']), 'ms/pass', round(d.get('ms_per_pass'),5), 'by rank', [round(x,5) for x in d['ms_per_pass_by_rank']], 'halo_ms', d.get('halo_ms_per_pass'), 'nccl', d.get('nccl_halo_ms_per_pass'), 'frac', round(d['roofline']['frac'],3), 'clk', d['clocks'].get('sm_mhz'), d['clocks'].get('reasons'), 'e2e', (d.get('e2e') or {}).get('value'), (d.get('e2e') or {}).get('pcie_h2d_GBps_plain_memcpy'), 'e2e_u8', (d.get('e2e_u8') or {}).get('value'), d['config']['output_checksum'])
except Exception as e: print(sys.argv[1], 'ERR', e)
PY
done
for f in $O/r2_scale_n*.err; do tail -n 2 $f; done
