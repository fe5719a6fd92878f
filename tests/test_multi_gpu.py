"""One process per GPU over NCCL (the deployment of BASELINE.json's 1/2/4/8-GPU runs): the sharded run reproduces the
single-GPU run -- AMM-PGO# bit for bit, AMM-PGO* to 1e-9 -- with the library's own NCCL transport and with the
torch.distributed callbacks.  Needs at least two devices; runs tools/multi_gpu_check.py under torchrun."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4])
def test_sharded_over_nccl_reproduces_single_gpu(world):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr",
           "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "multi_gpu_check.py")]
    # the script carries its own watchdog (a hang dumps the Python stacks and exits after 240 s per case)
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=1500)
    assert "MULTI_GPU_CHECK PASS" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
