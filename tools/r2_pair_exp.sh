#!/usr/bin/env bash
# A/B of the frame-pair front-end kernel (tcr_mfcc_pair.cu): the library in the tree against a previous build kept under
# tools/ab/ (not in history) and against the one-frame-per-warp kernel (TCR_MFCC_PAIR=0), on one box.
set -u
O=gpurun_out/${1:-pair3}; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_golden.py -m gpu -x -q > $O/pytest_pair.txt 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_pair.txt
show() {
python - "$1" "$2" <<'P'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    ks = " ".join(f"{k['name']}:{k['us']:.0f}" for k in d["kernels"][:6])
    print(sys.argv[1], round(d["value"]), round(d["ms_per_step"], 4), ks)
except Exception as e:
    print(sys.argv[1], "failed", e)
P
}
run() {  # label frontend env...
  local label=$1 fe=$2; shift 2
  env "$@" timeout 200 python bench.py --steps 100 --warmup 10 --frontend $fe --no-cpu-baseline --no-e2e --no-extra > $O/b_$label.json 2> $O/b_$label.err
  show "$label" $O/b_$label.json
}
run old_ordered ordered TCR_MFCC_PAIR=0
run new_ordered ordered TCR_MFCC_PAIR=1
run new_ahead ahead TCR_MFCC_PAIR=1
if [ -f tools/ab/libtcr_pair_v2.so ]; then
  cp tc-resnet_b200/libtcr_b200.so $O/new.so; cp tools/ab/libtcr_pair_v2.so tc-resnet_b200/libtcr_b200.so
  run prev_ordered ordered TCR_MFCC_PAIR=1
  run prev_ahead ahead TCR_MFCC_PAIR=1
  cp $O/new.so tc-resnet_b200/libtcr_b200.so; rm -f $O/new.so
fi
run new_ordered2 ordered TCR_MFCC_PAIR=1
TCR_MFCC_PAIR=1 timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mfcc -c 2 -o $O/mfcc_pair \
  python bench.py --steps 2 --warmup 1 --frontend ordered --no-cpu-baseline --no-e2e --no-extra > $O/ncu.log 2>&1; echo "ncu rc=$?"
