#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-resprof}; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:resident -s 6 -c 2 -o $O/res python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/ncu.log
