#!/usr/bin/env bash
set -u
O=gpurun_out/${1:-dsprof}; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:dsblock_ws -s 4 -c 1 -o $O/ds python bench.py --workload dscnn --steps 2 --warmup 2 > $O/ncu.log 2>&1; echo "ncu rc=$?"
