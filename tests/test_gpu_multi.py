"""Multi-GPU parity on the device (run with -m gpu on a box with >= 2 GPUs; skipped on a single-GPU box): the row-sharded
NCCL path of csrc/cg.cu / comm.cu -- all-gather of the direction block, all-reduce of the packed dot products -- must give
the single-GPU result and match the oracle (tools/mgpu_check.py --assert, one process per GPU under torch.distributed.run)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.timeout(900)
@pytest.mark.parametrize("world", [2, 4])
def test_row_sharded_nccl_path_matches_single_gpu_and_oracle(cuda_dev, world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    port = 29500 + world + (os.getpid() % 400)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tools", "mgpu_check.py"), "--assert"]
    env = dict(os.environ)
    env.setdefault("NCCL_DEBUG", "WARN")
    r = subprocess.run(cmd, cwd=ROOT, env=env, capture_output=True, text=True, timeout=850)
    assert r.returncode == 0, f"stdout:\n{r.stdout[-4000:]}\nstderr:\n{r.stderr[-6000:]}"
    assert f"mgpu_check ok (world={world}, assert=True)" in r.stdout
