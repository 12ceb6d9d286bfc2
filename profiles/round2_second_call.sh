#!/bin/bash
# Second GPU call of round 2 (one B200, ~3 min): in-engine sweep of the two knobs of the prefiltered read phase that were fixed from
# the microbenchmark only — the key block size (VB_KEY_BLOCK_MB, default 52) and the persisting-L2 set-aside (VB_BLOCK_L2_MB, default 64).
#   /usr/local/graft/bin/gpurun --timeout 400 -- 'bash profiles/round2_second_call.sh'
mkdir -p gpurun_out
for kb in 36 52 70 105; do
  for l2 in 64 79; do
    (VB_KEY_BLOCK_MB=$kb VB_BLOCK_L2_MB=$l2 timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu) > gpurun_out/bench_kb${kb}_l2${l2}.json 2> gpurun_out/bench_kb${kb}_l2${l2}.err
    python -c "import json; d=json.load(open('gpurun_out/bench_kb${kb}_l2${l2}.json')); print('key block $kb MB, L2 set-aside $l2 MB:', round(d['ms_per_step'],3), 'ms/step, kernels', round(d['roofline']['kernel_ms'],3), 'ms,', d['roofline']['kernel'][:60])"
  done
done
