#!/bin/bash
# Round-2 ncu evidence on one B200 (numbers printed by runs under ncu are never bench values):
#   (1) launch list of the read phase of one apply! with DRAM bytes per launch  -> r2_launches_hk_segsweep.csv (feeds hk_step_traffic.json)
#   (2) --set full of the two segmented sweeps                                   -> r2_hk_segsweep.ncu-rep
#   (3) --set full of the strip-shaped grid-stencil kernel (Game of Life 4096^2) -> r2_gol_strip.ncu-rep
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none \
   -k regex:"build_keys|reduce_segsweep|reduce_hubmerge|passrate" -c 24 --csv --log-file gpurun_out/r2_launches_hk_segsweep.csv \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_segsweep -s 4 -c 2 -o gpurun_out/r2_hk_segsweep \
   python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:reduce_stencil_strip -s 5 -c 1 -o gpurun_out/r2_gol_strip \
   python profiles/bench_configs.py gol 1.0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/r2_launches_hk_segsweep.csv
