#!/bin/bash
# Multi-GPU parity on hardware (round 2):  gpurun --gpus N --timeout 900 -- 'bash profiles/r2_mgpu_tests.sh'
# Runs the multi-rank pytest files on all visible GPUs (<= 4 ranks per script), then the sharded HK parity script with the
# blocking / prefilter thresholds lowered, so the prefiltered [local | ghost] sweeps that the scaling bench times are compared with
# the single-rank oracle.  The log is copied to profiles/r2_mgpu_tests_<N>gpu.txt.
mkdir -p gpurun_out
N=$(python -c "import torch; print(torch.cuda.device_count())")
OUT=gpurun_out/r2_mgpu_tests_${N}gpu.txt
{
  echo "# $(nvidia-smi -L | wc -l) GPUs visible; $(date -u +%FT%TZ)"
  timeout 1500 python -m pytest tests/test_multigpu.py tests/test_zzz_mgpu_next.py -m gpu -v -rs --tb=short 2>&1 | grep -v "^W1\|^\*\*\*\|OMP_NUM" | tail -n 150
  echo "# sharded HK, prefiltered sweeps forced on the parity graph (VB_BLOCK_EAGER=1 VB_BLOCK_MIN_MB=0 VB_KEY_BLOCK_MB=0.05)"
  MGPU_EXPECT_PREFILTER=1 VB_BLOCK_EAGER=1 VB_BLOCK_MIN_MB=0 VB_KEY_BLOCK_MB=0.05 VB_HALO_PHASES=3 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $((N<4?N:4)) --master-addr 127.0.0.1 --master-port 29529 tests/mgpu_hk.py 2>&1 | grep -v "^W\|^\*\*\*" | tail -n 30
} > $OUT 2>&1
cat $OUT
