#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-mfsweep}; mkdir -p $O
for cfg in 7,7 14,7 10,5 17,6 25,7 13,7 7,4; do
  TCR_MFCC_FPB=$cfg timeout 300 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-e2e --no-extra > $O/b_$cfg.json 2> $O/b_$cfg.err
  python - <<P
import json
try:
    d=json.load(open('$O/b_$cfg.json'))
    k=[x for x in d['kernels'] if x['name']=='mfcc'][0]
    print('$cfg', round(d['ms_per_step'],4), 'mfcc', round(k['us'],1))
except Exception as e: print('$cfg failed', e)
P
done
