#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-ahead}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_augment.py tests/test_reference_api.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt
for flag in "" "--ordered-frontend"; do
  timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-extra $flag > $O/bench$flag.json 2> $O/bench$flag.err; echo "bench '$flag' rc=$?"
  python - <<P
import json
d=json.load(open('$O/bench$flag.json'))
print('$flag', d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e'].get('pcm16_device_input_stage',{}).get('value'), d['e2e'].get('pcm16',{}).get('value'))
for k in d.get('kernels',[])[:7]: print('   ', k['name'], round(k['us'],1), round(k['share'],3))
P
done
