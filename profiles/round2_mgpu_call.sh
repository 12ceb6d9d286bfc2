#!/bin/bash
# First 2-GPU call of round 2 (~3 min of box time x 2 GPUs): the multi-rank paths written after round 1's GPU budget was spent.
#   /usr/local/graft/bin/gpurun --gpus 2 --timeout 420 -- 'bash profiles/round2_mgpu_call.sh'
mkdir -p gpurun_out
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port "$1" "$2"; }
(timeout 120 bash -c "$(declare -f run); run 29525 tests/mgpu_remove.py") > gpurun_out/mgpu_remove.txt 2>&1          # remote remove_edges! (removeedges_alltoall!)
(timeout 120 bash -c "$(declare -f run); run 29527 tests/mgpu_distribute.py") > gpurun_out/mgpu_distribute.txt 2>&1  # finish_init!(distribute = true)
(timeout 120 bash -c "$(declare -f run); run 29535 tests/mgpu_core.py") > gpurun_out/mgpu_core.txt 2>&1              # test/core.jl under mpiexec
(timeout 120 bash -c "$(declare -f run); run 29537 tests/mgpu_agentstate.py") > gpurun_out/mgpu_agentstate.txt 2>&1  # test/mpi/test_agentstate.jl
# the prefiltered sweeps on two ranks (keys of [local | ghost] slots) against the single-rank oracle: thresholds lowered so that
# the 200k-agent parity graph takes the swept read phase
(VB_BLOCK_EAGER=1 VB_BLOCK_MIN_MB=0 VB_KEY_BLOCK_MB=0.05 timeout 120 bash -c "$(declare -f run); run 29529 tests/mgpu_hk.py") > gpurun_out/mgpu_hk_prefilter.txt 2>&1
(VB_TRACE=1 timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu) > gpurun_out/bench_2gpu_prefilter.json 2> gpurun_out/bench_2gpu_prefilter.err
tail -n 4 gpurun_out/mgpu_remove.txt gpurun_out/mgpu_distribute.txt gpurun_out/mgpu_core.txt gpurun_out/mgpu_agentstate.txt gpurun_out/mgpu_hk_prefilter.txt
tail -n 1 gpurun_out/bench_2gpu_prefilter.json
