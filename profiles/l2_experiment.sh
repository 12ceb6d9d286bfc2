#!/bin/bash
# L2 residency experiment for the HK read phase: VB_L2_PERSIST_MB = size of the persisting window over the head of the state array
for mb in 0 48 80; do
  VB_L2_PERSIST_MB=$mb timeout 300 python bench.py --steps 10 --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]);print('persist_mb',$mb,'ms_per_step',round(d['ms_per_step'],3),'kernel_ms',round(d['roofline']['kernel_ms'],3),'frac',round(d['roofline']['frac'],4))"
done
