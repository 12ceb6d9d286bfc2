"""Multi-GPU tests of the paths added after the first multi-rank suite (remote remove_edges!, finish_init!(distribute = true), the
reference's core / agentstate tests under several ranks, rasters handed out to the ranks).  Green on 2 and 4 B200
(profiles/r2_mgpu_tests_{2,4}gpu.txt)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
def test_remove_edges_across_ranks(cuda):
    """removeedges_alltoall! (src/MPI.jl:432-479): """
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_remove.py", 29525)


@pytest.mark.gpu
def test_finish_init_distribute_across_ranks(cuda):
    """finish_init!(distribute = true) (src/MPI.jl:11-84): """
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_distribute.py", 29527)


@pytest.mark.gpu
def test_core_jl_across_ranks(cuda):
    """test/core.jl under mpiexec (test/mpi/test_core.jl): """
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_core.py", 29535)


@pytest.mark.gpu
def test_agentstate_jl_across_ranks(cuda):
    """test/mpi/test_agentstate.jl: """
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_agentstate.py", 29537)


@pytest.mark.gpu
def test_raster_handed_out_game_of_life_across_ranks(cuda):
    """broadcastids + join (src/MPI.jl:59-73,492-517; src/Raster.jl:64-75,227,318,378): the cells of a raster are handed out, rastervalues /
    calc_raster join the ranks; Game of Life on 2-4 ranks against numpy and the single-rank oracle"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_gol.py", 29539, nranks=2)          # the configuration that ran on hardware (profiles/r2_mgpu_tests_2gpu.txt)


@pytest.mark.gpu
def test_agents_placed_on_a_handed_out_raster_across_ranks(cuda):
    """move_to! in the initialisation phase and inside a transition with cells of other ranks (src/Raster.jl:437-477; test/raster.jl:242-303
    as mpiexec runs it): the staged edges are handed out with their targets, Ctx::move_to towards a remote cell sends both edges through
    transmit_edges!, the joined calc_rasterstate follows the single-rank oracle"""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    from mgpu_common import run_ranks
    run_ranks("mgpu_moveto.py", 29541, nranks=2)       # the configuration that ran on hardware (profiles/r2_mgpu_moveto_2gpu.txt)
