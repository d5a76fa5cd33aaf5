#!/usr/bin/env bash
set -u
N=${1:-2}; O=gpurun_out/${2:-multi$N}; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 100 --warmup 10 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?"; tail -3 $O/bench_n$N.err
python - <<P
import json
d=json.load(open('$O/bench_n$N.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e', d.get('e2e',{}).get('value')); print('dp_check', d.get('dp_check')); print(d.get('config'))
for k in d.get('kernels',[])[:6]: print(k['name'], round(k['us'],1), round(k['share'],3))
P
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $O/ref_n$N.json 2> $O/ref_n$N.err; echo "ref N=$N rc=$?"; cut -c1-400 $O/ref_n$N.json
timeout 900 python -m pytest tests/test_dist.py -m gpu -x -q > $O/pytest_dist.txt 2>&1; echo "pytest dist rc=$?"; tail -3 $O/pytest_dist.txt
