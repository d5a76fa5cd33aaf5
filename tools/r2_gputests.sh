#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-gt}; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
tail -25 $O/pytest_gpu.txt
