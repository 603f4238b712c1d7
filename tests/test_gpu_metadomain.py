"""Multi-domain exchange on the GPU(s): the CUDA library's pack / NCCL / unpack path against
the single-domain oracle and the numpy restatement of the protocol (tests/mgpu_worker.py).
With one visible GPU the same code runs with one domain that is its own periodic neighbour
(generic pack -> self copy -> unpack path); with >= 2 GPUs it is launched under torchrun."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("classify", ["block-sort", "device-sort", "scan"])
def test_single_domain_through_generic_exchange(classify):
    """classify: how the migration finds the dead / outgoing particles -- from the fused pusher's
    exception list (sorted by one block, or by the device-wide radix sort for long lists) or by
    scanning every tag (EB200_MIGRATE_SCAN=1, the path of kernels that keep no list)"""
    if _ngpu() < 1:
        pytest.skip("no CUDA device")
    env = dict(os.environ, WORLD_SIZE="1", RANK="0", LOCAL_RANK="0")
    env.pop("EB200_EXC_SMALL", None)
    env.pop("EB200_MIGRATE_SCAN", None)
    if classify == "device-sort":
        env["EB200_EXC_SMALL"] = "0"
    elif classify == "scan":
        env["EB200_MIGRATE_SCAN"] = "1"
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "mgpu_worker.py")], env=env,
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]


@pytest.mark.parametrize("world,d2,d3", [(2, "-1,2", "-1,-1,-1"), (2, "2,1", "1,2,1"),
                                         (4, "-1,2", "-1,-1,-1"), (8, "-1,2", "-1,-1,-1")])
def test_domains_over_nccl(world, d2, d3):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, EB200_DECOMP2D=d2, EB200_DECOMP3D=d3)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
           os.path.join(ROOT, "tests", "mgpu_worker.py")]
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=1500)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
