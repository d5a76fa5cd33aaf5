#!/usr/bin/env bash
set -u
N=${1:-2}; O=gpurun_out/${2:-multiab$N}; mkdir -p $O
for fe in ordered ahead; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 --frontend $fe --no-e2e > $O/bench_n${N}_$fe.json 2> $O/bench_n${N}_$fe.err; echo "bench N=$N $fe rc=$?"
python - <<P
import json
d=json.load(open('$O/bench_n${N}_$fe.json'))
print('$fe', d['value'], d['ms_per_step'], d.get('dp_check',{}).get('ok'))
for k in d.get('kernels',[])[:7]: print('   ', k['name'], round(k['us'],1), round(k['share'],3))
P
done
