"""Context-parallel DiT on >= 2 GPUs: sharded forward == single-GPU forward (tools/cp_check.py under torchrun)."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4, 8])
def test_context_parallel_matches_single_gpu(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29500 + world),
                        os.path.join(ROOT, "tools", "cp_check.py")], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "CP_CHECK_PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("world", [2, 4])
def test_vae_units_sharded_over_ranks_match_single_gpu(world):
    """decode_latent chunks / decode_tiled tiles round-robin over ranks (tools/vae_cp_check.py under torchrun)."""
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
                        "--master-addr", "127.0.0.1", "--master-port", str(29520 + world),
                        os.path.join(ROOT, "tools", "vae_cp_check.py")], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and "VAE_CP_CHECK_PASS" in r.stdout, (r.stdout[-3000:], r.stderr[-3000:])
