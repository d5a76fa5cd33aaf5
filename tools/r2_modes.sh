#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-modes}; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "resident or full_size or bitwise" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest.txt
for m in 3 1; do
  TCR_RESIDENT=$m timeout 600 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e --no-extra > $O/bench_m$m.json 2> $O/bench_m$m.err; echo "bench mode $m rc=$?"
  python - <<P
import json
d=json.load(open('$O/bench_m$m.json'))
print('mode $m', d['value'], d['ms_per_step'])
for k in d.get('kernels',[])[:8]: print('   ', k['name'], round(k['us'],1), round(k['share'],3))
P
done
TCR_RESIDENT=3 timeout 600 python bench.py --model TCResNet14 --width 1.5 --batch 1024 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e --no-extra > $O/bench14_m3.json 2> $O/bench14_m3.err; echo "r14 mode3 rc=$?"
python -c "
import json; d=json.load(open('$O/bench14_m3.json')); print('r14 m3', d['value'], d['ms_per_step']); [print('   ',k['name'], round(k['us'],1)) for k in d['kernels'][:6]]"
