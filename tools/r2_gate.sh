#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-gate}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "frontend or bitwise or resident" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -2 $O/pytest.txt
for fe in ordered ahead ordered ahead; do
  timeout 300 python bench.py --steps 300 --warmup 20 --no-cpu-baseline --no-extra --no-e2e --frontend $fe > $O/b_$fe.json 2> $O/b_$fe.err
  python -c "
import json; d=json.load(open('$O/b_$fe.json')); print('$fe', round(d['value']), round(d['ms_per_step'],4), ' '.join(k['name']+':'+str(round(k['us'])) for k in d['kernels'][:6]))"
done
