"""Multi-GPU tests written after round 1's GPU budget was spent (never run on GPUs yet): they sort behind the established suite so
that `pytest -x` reaches them last."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_remove_edges_across_ranks(cuda):
    """removeedges_alltoall! (src/MPI.jl:432-479): written after round 1's GPU budget was spent, not run on GPUs yet"""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(ng, 4)), "--master-addr", "127.0.0.1",
                        "--master-port", "29525", os.path.join(ROOT, "tests", "mgpu_remove.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": ok") == min(ng, 4)


@pytest.mark.gpu
def test_finish_init_distribute_across_ranks(cuda):
    """finish_init!(distribute = true) (src/MPI.jl:11-84): written after round 1's GPU budget was spent, not run on GPUs yet"""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(ng, 4)), "--master-addr", "127.0.0.1",
                        "--master-port", "29527", os.path.join(ROOT, "tests", "mgpu_distribute.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": ok") == min(ng, 4)


@pytest.mark.gpu
def test_core_jl_across_ranks(cuda):
    """test/core.jl under mpiexec (test/mpi/test_core.jl): written after round 1's GPU budget was spent, not run on GPUs yet"""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(ng, 4)), "--master-addr", "127.0.0.1",
                        "--master-port", "29535", os.path.join(ROOT, "tests", "mgpu_core.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": ok") == min(ng, 4)


@pytest.mark.gpu
def test_agentstate_jl_across_ranks(cuda):
    """test/mpi/test_agentstate.jl: written after round 1's GPU budget was spent, not run on GPUs yet"""
    import torch
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(min(ng, 4)), "--master-addr", "127.0.0.1",
                        "--master-port", "29537", os.path.join(ROOT, "tests", "mgpu_agentstate.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count(": ok") == min(ng, 4)
