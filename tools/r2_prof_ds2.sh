#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-dsprof}; mkdir -p $O
timeout 180 python -m pytest tests/test_dscnn.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest(tc) rc=$?"; tail -2 $O/pytest.txt
timeout 120 python bench.py --workload dscnn --steps 100 --warmup 10 > $O/bench_tc1.json 2> $O/bench_tc1.err; echo "bench rc=$?"
python -c "import json; d=json.load(open('$O/bench_tc1.json')); print(d['value'], 'utt/s', d['ms_per_step'], 'ms')"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dsblock_ws -s 4 -c 1 -o $O/ds python bench.py --workload dscnn --steps 2 --warmup 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:conv_ws -s 2 -c 1 -o $O/conv python bench.py --workload dscnn --steps 2 --warmup 2 > $O/ncu2.log 2>&1; echo "ncu rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none --kernel-name-base demangled -k regex:dscnn -s 12 -c 6 --csv --log-file $O/ncu_dscnn.csv python bench.py --workload dscnn --steps 3 --warmup 3 > $O/ncu3.log 2>&1; echo "ncu rc=$?"; grep time_duration $O/ncu_dscnn.csv | cut -d, -f5,12- | cut -c1-150
