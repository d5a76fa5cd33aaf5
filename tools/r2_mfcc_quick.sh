#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-mfq}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt
timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extra > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
python - <<P
import json
d=json.load(open('$O/bench.json'))
print(d['value'], d['ms_per_step'])
for k in d.get('kernels',[])[:14]: print(k['name'], round(k['us'],1), round(k['share'],3))
P
