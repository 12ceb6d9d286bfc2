#!/bin/bash
# short bench on N ranks (default: all visible GPUs), optional device-time marks with TRACE=2
N=${1:-$(python -c "import torch; print(torch.cuda.device_count())")}
mkdir -p gpurun_out
VB_TRACE=${TRACE:-} timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$((N%10)) bench.py --gpus $N --steps ${STEPS:-10} --warmup 3 --no-cpu 2> gpurun_out/bench_${N}gpu.err | tail -1 > gpurun_out/bench_${N}gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_${N}gpu.json')); print('N=$N ms/step', d['ms_per_step'], 'opinion_sum', d['config']['opinion_sum'], 'kernel_ms', d['roofline']['kernel_ms'])"
if [ -n "$TRACE" ]; then
  grep "vb mark\|vb halo" gpurun_out/bench_${N}gpu.err | tail -$((N*26)) | grep "sweep of\|push of\|keys of\|barrier" | sed 's/: [0-9]* entries//; s/: [0-9]* B//; s/(at.*//' | awk '{k=$3" "$4" "$5" "$6" "$7; v=$(NF-1); s[k]+=v; c[k]++; if(v>m[k])m[k]=v} END{for(k in s) printf "   %-40s mean %.3f max %.3f (n=%d)\n", k, s[k]/c[k], m[k], c[k]}' | sort
  grep "last mark" gpurun_out/bench_${N}gpu.err | tail -$N | awk '{print $(NF-1)}' | sort -n | tail -1 | sed 's/^/   main stream ends at (max) /'
fi
