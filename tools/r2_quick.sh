#!/usr/bin/env bash
# quick GPU iteration: parity tests, resident-kernel timelines, short bench (per-kernel table), TCR_RESIDENT=1 / 0 comparison.
set -u
T=${1:-q}; O=gpurun_out/$T; mkdir -p $O
TCR_RESIDENT_VERBOSE=1 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > $O/pytest_gpu.txt 2>&1; echo "pytest rc=$?" | tee -a $O/pytest_gpu.txt
grep -m12 "\[tcr\]" $O/pytest_gpu.txt; tail -8 $O/pytest_gpu.txt
timeout 300 python tools/timeline_res.py > $O/timeline.txt 2>&1; echo "timeline rc=$?"; tail -75 $O/timeline.txt
for mode in 2 1 0; do
  TCR_RESIDENT=$mode timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e > $O/bench_m$mode.json 2> $O/bench_m$mode.err; echo "bench(mode $mode) rc=$?"; tail -2 $O/bench_m$mode.err
  python - <<PY
import json
d=json.load(open("$O/bench_m$mode.json"))
print("mode $mode ms_per_step", d["ms_per_step"], "value", d["value"], "loss", d.get("final_total_loss"))
for k in d.get("kernels",[])[:8]: print(f'  {k["name"]:28s} {k["us"]:8.2f} us  share {k["share"]:.3f}')
PY
done
