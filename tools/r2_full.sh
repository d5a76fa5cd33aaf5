#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-full}; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest_gpu.txt
timeout 600 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; tail -3 $O/bench.err
python - <<P
import json
d=json.load(open('$O/bench.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches') if k in d}); print(d.get('e2e')); print(d.get('roofline',{}).get('kernel'), d.get('roofline',{}).get('frac')); print(d.get('clocks'))
for k in d.get('kernels',[])[:10]: print(k['name'], round(k['us'],1), round(k['share'],3))
print(json.dumps(d.get('extra'))[:1500])
P
timeout 300 python bench.py --workload dscnn --steps 100 --warmup 10 > $O/bench_dscnn.json 2> $O/bench_dscnn.err; echo "dscnn rc=$?"; tail -2 $O/bench_dscnn.err
python -c "
import json; d=json.load(open('$O/bench_dscnn.json')); print(d['value'], d['ms_per_step'], d['roofline']['frac'], d['e2e']['value']); [print(k['name'], round(k['us'],1)) for k in d['kernels']]"
