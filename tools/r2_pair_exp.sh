#!/usr/bin/env bash
# A/B of the frame-pair front-end kernel (TCR_MFCC_PAIR, tcr_mfcc_pair.cu) against the one-frame-per-warp kernel:
# GPU parity tests with the pair kernel on, short bench runs in both front-end modes, one ncu --set full capture per variant.
set -u
O=gpurun_out/${1:-pair2}; mkdir -p $O
TCR_MFCC_PAIR=1 TCR_MFCC_PAIR_VARIANT=${2:-0} timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_augment.py tests/test_reference_api.py -m gpu -x -q > $O/pytest_pair.txt 2>&1; echo "pytest(pair) rc=$?"; tail -3 $O/pytest_pair.txt
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
for run in ordered:0:0 ordered:1:0 ordered:1:1 ordered:8:0 ahead:1:0 ahead:1:1; do
  IFS=: read fe cfg var <<< "$run"
  TCR_MFCC_PAIR=$cfg TCR_MFCC_PAIR_VARIANT=$var timeout 200 python bench.py --steps 100 --warmup 10 --frontend $fe --no-cpu-baseline --no-e2e --no-extra > $O/b_${fe}_${cfg}_$var.json 2> $O/b_${fe}_${cfg}_$var.err
  show "$fe pair=$cfg variant=$var" $O/b_${fe}_${cfg}_$var.json
done
for var in 0 1; do
  TCR_MFCC_PAIR=1 TCR_MFCC_PAIR_VARIANT=$var timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mfcc -c 2 -o $O/mfcc_pair_v$var \
    python bench.py --steps 2 --warmup 1 --frontend ordered --no-cpu-baseline --no-e2e --no-extra > $O/ncu_v$var.log 2>&1; echo "ncu v$var rc=$?"
done
