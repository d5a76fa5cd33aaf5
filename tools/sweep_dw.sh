#!/usr/bin/env bash
# Sweep the weight-gradient chunking knobs (TCR_DW_MACS = MACs per CTA, TCR_DW_UPC = max utterances per chunk).
for cfg in "0.5e6 8" "1e6 8" "1e6 16" "2e6 16" "2e6 32" "4e6 32" "0.25e6 4"; do
  set -- $cfg
  TCR_DW_MACS=$1 TCR_DW_UPC=$2 python bench.py --steps 200 --warmup 10 --no-cpu-baseline --no-e2e "${@:3}" > /tmp/sw.json 2>/dev/null
  python - <<PY
import json
d=json.load(open("/tmp/sw.json"))
k={x["name"]:x["us"] for x in d["kernels"]}
print("macs=$1 upc=$2", round(d["ms_per_step"],4), "dw", round(k.get("dw_grouped",0),1), "gradfin", round(k.get("grad_finalize",0),1))
PY
done
