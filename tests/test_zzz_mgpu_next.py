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
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_remove.py", 29525)


@pytest.mark.gpu
def test_finish_init_distribute_across_ranks(cuda):
    """finish_init!(distribute = true) (src/MPI.jl:11-84): written after round 1's GPU budget was spent, not run on GPUs yet"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_distribute.py", 29527)


@pytest.mark.gpu
def test_core_jl_across_ranks(cuda):
    """test/core.jl under mpiexec (test/mpi/test_core.jl): written after round 1's GPU budget was spent, not run on GPUs yet"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_core.py", 29535)


@pytest.mark.gpu
def test_agentstate_jl_across_ranks(cuda):
    """test/mpi/test_agentstate.jl: written after round 1's GPU budget was spent, not run on GPUs yet"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_agentstate.py", 29537)
