#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-mfccprof}; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_reference_api.py -m gpu -x -q -k "different_size or deployable or trainer_loop" > $O/pytest.txt 2>&1; echo "pytest rc=$?"; tail -5 $O/pytest.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:mfcc_kernel -s 3 -c 1 -o $O/mfcc python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $O/ncu.log 2>&1; echo "ncu rc=$?"
