#!/bin/bash
# First GPU call of round 2 (one B200, ~2 min): the measurements that were prepared after round 1's GPU budget was spent.
#   /usr/local/graft/bin/gpurun --timeout 300 -- 'bash profiles/round2_first_call.sh'
# (the two microbenchmark binaries must have been compiled in the container first:
#   cd profiles/microbench && for f in wavefront prefilter stencil; do nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o $f $f.cu; done)
mkdir -p gpurun_out
cd profiles/microbench
(timeout 60 ./wavefront) > ../../gpurun_out/wavefront.txt 2>&1                                       # divergent byte-gather ceiling: LDG / L1 / texture / shared memory
(timeout 60 ./stencil 4096 4096) > ../../gpurun_out/stencil.txt 2>&1                                  # grid-stencil shapes for Game of Life (config 2)
(SHAPES2=1 POWERLAW=1 timeout 60 ./prefilter 100000000 20) > ../../gpurun_out/prefilter_shapes3.txt 2>&1   # CTA sizes, texture keys, mask row find
cd ../..
for w in 2 4 8; do                                                                                    # in-engine time of the sweep shapes (default 2)
  (VB_PF_WARPS=$w timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu) > gpurun_out/bench_pf_warps$w.json 2> gpurun_out/bench_pf_warps$w.err
done
# experimental mask row find in the engine kernel: parity first, then time
(VB_PF_ROWFIND=1 timeout 100 python -m pytest tests/test_hk.py -x -q -m gpu -k "not full_size") > gpurun_out/test_hk_rowfind.txt 2>&1
(VB_PF_ROWFIND=1 timeout 100 python bench.py --steps 5 --warmup 3 --no-cpu) > gpurun_out/bench_pf_rowfind.json 2> gpurun_out/bench_pf_rowfind.err
# GPU variants written after round 1's GPU budget was spent (market model, registered maps, golden replays, full-size cases)
(timeout 400 python -m pytest tests/test_zzm_market.py tests/test_zzn_remove_order.py tests/test_zzo_edge_cases.py tests/test_zzp_agentstate.py tests/test_zzq_raster_maps.py tests/test_zzr_ext_model_gpu.py tests/test_zw_golden.py tests/test_zx_misc.py tests/test_zy_full_size.py tests/test_zz_long_runs.py -q -m gpu) > gpurun_out/test_gpu_late.txt 2>&1
tail -n 5 gpurun_out/test_gpu_late.txt
tail -n 3 gpurun_out/test_hk_rowfind.txt
cat gpurun_out/stencil.txt
tail -n 40 gpurun_out/wavefront.txt gpurun_out/prefilter_shapes3.txt
for w in 2 4 8; do python -c "import json,sys; d=json.load(open('gpurun_out/bench_pf_warps$w.json')); print('VB_PF_WARPS=$w', d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; done
