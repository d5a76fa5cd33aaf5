#!/usr/bin/env bash
# round-2 baseline: tests, bench, launch list and one --set full capture of the step's kernels at HEAD
set -u
O=gpurun_out/r2_base; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1; nproc >> $O/smi.txt
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
tail -5 $O/pytest_gpu.txt
timeout 600 python bench.py --steps 100 --warmup 10 > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"
cut -c1-1500 $O/bench.json
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tcr:: -s 40 -c 21 -o $O/full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
