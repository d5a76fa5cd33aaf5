#!/usr/bin/env bash
# Sweep the weight-gradient chunking knobs (TCR_DW_SLOTS = CTA budget of the balanced plan, TCR_DW_UPC = max utterances per chunk).
for cfg in "296 64" "280 64" "296 32" "444 64" "592 64" "148 64"; do
  set -- $cfg
  TCR_DW_SLOTS=$1 TCR_DW_UPC=$2 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e "${@:3}" > /tmp/sw.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("/tmp/sw.json"))
k={x["name"]:x["us"] for x in d["kernels"]}
print("slots=$1 upc=$2", round(d["ms_per_step"],4), "dw", round(k.get("dw_grouped",0),1), "gradfin", round(k.get("grad_finalize",0),1))
PY
done
