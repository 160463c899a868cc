"""The partitioned (N>1) path: world_size-2 gloo run on CPU for the host plumbing, torchrun + NCCL on >= 2 GPUs for parity."""
import os
import subprocess
import sys

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _run(world, mode, case, port):
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "multi_rank_worker.py"), mode, case]
    env = dict(os.environ, OMP_NUM_THREADS="2")
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900, env=env)
    assert p.returncode == 0 and "MULTI_RANK_OK" in p.stdout, p.stdout[-3000:] + p.stderr[-3000:]


@pytest.mark.parametrize("world,case", [(2, "split"), (3, "weakG"), (2, "config4"), (3, "splitG"), (2, "splitCH"), (3, "config4_unstr"), (3, "naca"), (2, "splitfilt"), (3, "hopr"), (2, "n11")])
def test_partition_plumbing_gloo(world, case):
    _run(world, "cpu", case, 29611 + world)


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["split", "weakG", "euler", "config4", "splitwall", "splitG", "splitCH", "config4_unstr", "split_unstr", "naca", "splitfilt", "hopr", "n11"])
def test_nccl_side_exchange_parity(case):
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (gpurun --gpus 2)")
    _run(4 if n >= 4 else 2, "gpu", case, 29701)
