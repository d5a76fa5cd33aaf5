#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-dsq}; mkdir -p $O
timeout 180 python -m pytest tests/test_dscnn.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest(tc) rc=$?"; tail -2 $O/pytest.txt
timeout 120 python bench.py --workload dscnn --steps 100 --warmup 10 > $O/bench_tc1.json 2> $O/bench_tc1.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/bench_tc1.json')); print(d['value'], 'utt/s', d['ms_per_step'], 'ms')"
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum --clock-control none --kernel-name-base demangled -k regex:dscnn -s 12 -c 6 --csv --log-file $O/ncu_dscnn.csv python bench.py --workload dscnn --steps 3 --warmup 3 > $O/ncu3.log 2>&1; echo "ncu rc=$?"
python - <<P
import csv
rows=[r for r in csv.reader(open('$O/ncu_dscnn.csv')) if len(r)>5]
h=rows[0]; ki=h.index('Kernel Name'); mi=h.index('Metric Name'); vi=h.index('Metric Value')
for r in rows[1:]: print(r[ki][:36], r[mi][:40], r[vi])
P
