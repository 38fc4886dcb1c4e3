"""Slab decomposition across GPUs (needs >= 2 GPUs: `gpurun --gpus 2 -- python -m pytest tests/test_multigpu.py -m gpu`).
A lattice cut into slabs with NCCL halo exchange must evolve BIT-IDENTICALLY to the same lattice on one GPU, also at T > 0
(Philox is keyed by the site inside its plane and the GLOBAL plane index); energies agree to summation order."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_slabs_match_single_gpu(product, world):
    if product.SpiritB200_Device_Count() < world:
        pytest.skip("needs %d GPUs on the box" % world)
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world, "--master-addr", "127.0.0.1",
         "--master-port", "29714", os.path.join(ROOT, "tests", "mgpu_worker.py")],
        capture_output=True, text=True, timeout=900)
    sys.stdout.write(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "MGPU_FAILURES 0" in r.stdout
