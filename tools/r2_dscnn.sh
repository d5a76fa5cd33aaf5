#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-ds}; mkdir -p $O
timeout 600 python -m pytest tests/test_dscnn.py -m gpu -x -q > $O/pytest.txt 2>&1; echo "pytest(tc) rc=$?"; tail -4 $O/pytest.txt
TCR_DSCNN_TC=0 timeout 600 python -m pytest tests/test_dscnn.py -m gpu -x -q > $O/pytest_fma.txt 2>&1; echo "pytest(fma) rc=$?"; tail -2 $O/pytest_fma.txt
for tc in 1 0; do
  TCR_DSCNN_TC=$tc timeout 300 python bench.py --workload dscnn --steps 100 --warmup 10 > $O/bench_tc$tc.json 2> $O/bench_tc$tc.err; echo "bench tc=$tc rc=$?"; tail -2 $O/bench_tc$tc.err
  python -c "import json; d=json.load(open('$O/bench_tc$tc.json')); print('tc=$tc', d['value'], 'utt/s', d['ms_per_step'], 'ms', d.get('fp32_tflops'))"
done
timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --kernel-name-base demangled -k regex:dscnn -s 12 -c 6 --csv --log-file $O/ncu_dscnn.csv python bench.py --workload dscnn --steps 3 --warmup 3 > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -8 $O/ncu_dscnn.csv | cut -c1-300
