#!/usr/bin/env bash
# Final evidence of the round in one short gpurun call, most important first (the call may be cut by the GPU budget):
# GPU tests, ncu --set full of the training step (traffic table for bench.py's roofline), the default bench line, smoke(),
# the ncu launch list, racecheck of the front-end kernel.
set -u
O=gpurun_out/${1:-fin}; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv > $O/smi.txt 2>&1; nproc >> $O/smi.txt
timeout 300 python -m pytest tests -m gpu -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt; tail -3 $O/pytest_gpu.txt
timeout 200 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:tcr:: -s 12 -c 12 -o $O/full python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $O/ncu_full.log 2>&1; echo "ncu full rc=$?"
timeout 300 python bench.py > $O/bench.json 2> $O/bench.err; echo "bench rc=$?"; cut -c1-300 $O/bench.json
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $O/smoke.txt 2>&1; echo "smoke rc=$?"; tail -2 $O/smoke.txt
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $O/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-extra > $O/under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 150 compute-sanitizer --tool racecheck --error-exitcode 7 python -c "
import sys, numpy as np; sys.path.insert(0,'tests'); sys.path.insert(0,'.')
from tcr_harness import TorchBackend, Engine
from oracle import tcr_oracle as O
b = TorchBackend()
wav, _ = O.synthetic_batch(7, adversarial=True)
e = Engine(b, max_batch=8); f = e.mfcc(wav); print(f.shape, float(np.abs(f).max()))
" > $O/racecheck_mfcc.txt 2>&1; echo "racecheck rc=$?" | tee -a $O/racecheck_mfcc.txt; grep -E "RACECHECK SUMMARY|hazard" $O/racecheck_mfcc.txt | head -3
