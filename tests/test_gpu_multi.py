"""Multi-GPU parity (pytest -m gpu, skipped with fewer than 2 devices): torchrun over NCCL, the
slab run must be bit-identical to the single-GPU run (scripts/slab_check.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpus():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,overlap", [(2, 1), (2, 0), (4, 1)])
def test_slabs_bit_identical_to_single_gpu(world, overlap):
    if _ngpus() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), str(ROOT / "scripts" / "slab_check.py"),
           "1024", "1024", "3", "37"]
    # overlap 1: pressure exchanges overlapped with interior Jacobi launches (natrix_step_phase 4 / 5); 0: serial
    env = dict(os.environ, NATRIX_SLAB_OVERLAP=str(overlap))
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    assert res.returncode == 0 and "SLAB_CHECK PASS" in res.stdout, res.stdout[-3000:] + res.stderr[-3000:]


def test_pure_c_multi_gpu_host_is_bit_identical_to_single_gpu():
    """examples/c_host_multi.c: two host threads, two slab handles, natrix_comm_init + natrix_step - no Python and no
    torch in the process; every slab's velocity, pressure and dye rows equal the single-GPU run's byte for byte."""
    if _ngpus() < 2:
        pytest.skip("needs 2 GPUs")
    exe = ROOT / "examples" / "_build" / "c_host_multi"
    if not exe.exists():
        pytest.skip("examples/_build/c_host_multi is not built (python -c 'import __graft_entry__ as g; g.build()')")
    import torch
    nccl = os.path.join(os.path.dirname(os.path.dirname(torch.__file__)), "nvidia", "nccl", "lib", "libnccl.so.2")
    env = dict(os.environ)
    if os.path.exists(nccl):
        env["NATRIX_NCCL_LIB"] = nccl                      # the same NCCL the Python path uses
    res = subprocess.run([str(exe), "2", "6"], capture_output=True, text=True, timeout=300, env=env)
    assert res.returncode == 0 and "C_HOST_MULTI PASS" in res.stdout, res.stdout[-2000:] + res.stderr[-2000:]
