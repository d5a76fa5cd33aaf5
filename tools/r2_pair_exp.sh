#!/usr/bin/env bash
# A/B of the frame-pair front-end kernel (TCR_MFCC_PAIR, tcr_mfcc_pair.cu) against the one-frame-per-warp kernel:
# GPU parity suite with the pair kernel on, short bench runs in both front-end modes, one ncu --set full capture.
set -u
O=gpurun_out/${1:-pair1}; mkdir -p $O
TCR_MFCC_PAIR=1 timeout 400 python -m pytest tests -m gpu -x -q > $O/pytest_pair.txt 2>&1; echo "pytest(pair) rc=$?"; tail -3 $O/pytest_pair.txt
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
for run in ordered:0 ordered:1 ordered:8,4 ordered:10,3 ahead:0 ahead:1; do
  fe=${run%%:*}; cfg=${run##*:}
  {
    TCR_MFCC_PAIR=$cfg timeout 200 python bench.py --steps 100 --warmup 10 --frontend $fe --no-cpu-baseline --no-e2e --no-extra > $O/b_${fe}_$cfg.json 2> $O/b_${fe}_$cfg.err
    show "$fe pair=$cfg" $O/b_${fe}_$cfg.json
  }
done
TCR_MFCC_PAIR=1 timeout 300 ncu --set full --import-source on --clock-control none --kernel-name-base demangled -k regex:mfcc -c 2 -o $O/mfcc_pair \
  python bench.py --steps 2 --warmup 1 --frontend ordered --no-cpu-baseline --no-e2e --no-extra > $O/ncu.log 2>&1; echo "ncu rc=$?"
