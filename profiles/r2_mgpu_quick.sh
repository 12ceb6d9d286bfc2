#!/bin/bash
# quick multi-rank parity of the sharded HK step (direct + forced prefiltered [local | ghost] sweeps) and a short bench on all visible GPUs
N=$(python -c "import torch; print(torch.cuda.device_count())")
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tests/mgpu_hk.py 2>&1 | grep "rank\|Error\|error" | tail -12
for ph in 1 2 3; do
MGPU_EXPECT_PREFILTER=1 VB_BLOCK_EAGER=1 VB_BLOCK_MIN_MB=0 VB_KEY_BLOCK_MB=0.05 VB_HALO_PHASES=$ph timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29529 tests/mgpu_hk.py 2>&1 | grep "rank\|Error\|error" | tail -12
done
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu 2>/dev/null | tail -1 > gpurun_out/bench_${N}gpu_quick.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu_quick.json')); print('N=$N', d['ms_per_step'], d['config']['opinion_sum'], d['roofline']['kernel_ms'])"
