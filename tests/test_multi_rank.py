"""Multi-rank paths: world_size-2/3 gloo runs of the sharding logic on the CPU; NCCL runs on >= 2 GPUs (-m gpu)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WORKER = os.path.join(ROOT, "tests", "mg_worker.py")


def _launch(nproc, *args, timeout=300):
    port = 29500 + (os.getpid() % 400) + nproc
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), WORKER, *map(str, args)]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    assert r.returncode == 0 and "MG_OK" in r.stdout, (r.stdout[-2000:], r.stderr[-3000:])


@pytest.mark.parametrize("world,log2n,extra", [(2, 16, 0), (3, 16, 777)])
def test_sharded_halo_exchange_cpu_gloo(world, log2n, extra):
    _launch(world, "cpu", log2n, extra)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_halo_exchange_nccl(world):
    import sdr_b200
    if sdr_b200.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, "gpu", 24, 0)
    _launch(world, "gpu", 20, 4096 + 24)


@pytest.mark.gpu
@pytest.mark.parametrize("world", [2, 8])
def test_sharded_peer_memory_halo(world):
    """halo read in place from the neighbour's HBM (CUDA IPC + NVLink) instead of an NCCL exchange"""
    import sdr_b200
    if sdr_b200.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    _launch(world, "gpu-peer", 24, 0)
    _launch(world, "gpu-peer", 20, 4096 + 24)
