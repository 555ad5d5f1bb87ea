"""Data parallelism over NCCL on real GPUs (needs >= 2 of them: skipped on the driver's single-GPU test run, executed by
`gpurun --gpus 2 -- python -m pytest tests/test_dp_nccl_gpu.py -m gpu`).  tools/dp_check.py under torchrun:
  * the all-reduced gradients (early arena underneath the backbone dgrad, dSource exchanged in place of the mapping-layer
    gradient, late arena) equal the single-process gradient of the mean loss over the concatenated batch;
  * the captured-graph training step and the kernel-by-kernel step issue the same collective sequence and agree."""
import subprocess
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = Path(__file__).resolve().parent.parent


@pytest.mark.parametrize("world", [2, 4])
def test_dp_gradients_match_single_process_over_nccl(world, cuda):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, this box has {torch.cuda.device_count()}")
    port = 29500 + world
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(port), str(REPO / "tools" / "dp_check.py")],
                       capture_output=True, text=True, timeout=600, cwd=str(REPO))
    sys.stdout.write(r.stdout[-2000:])
    assert r.returncode == 0, r.stderr[-3000:]
    assert r.stdout.count("[dp_check]") >= 2
