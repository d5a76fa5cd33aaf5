#!/usr/bin/env bash
set -u
N=${1:-8}; O=gpurun_out/${2:-multi$N}; mkdir -p $O
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > $O/bench_n$N.json 2> $O/bench_n$N.err; echo "bench N=$N rc=$?"; tail -2 $O/bench_n$N.err
python - <<P
import json
d=json.load(open('$O/bench_n$N.json'))
print({k:d.get(k) for k in ('value','ms_per_step','n_gpus','gpu_launches')}); print('e2e', d.get('e2e',{}).get('value')); print('dp_check', d.get('dp_check')); print(d.get('config'))
for k in d.get('kernels',[])[:7]: print(k['name'], round(k['us'],1), round(k['share'],3))
P
