"""Print the headline numbers and the per-kernel table of a bench.py JSON line."""
import json, sys
d = json.load(open(sys.argv[1]))
print(*sys.argv[2:], f"{d['value']:.0f} {d['unit']}  {d['ms_per_step']:.4f} ms/step  launches/step {d['gpu_launches'] / d['steps']:.0f}")
if "e2e" in d:
    e = d["e2e"]
    print("  e2e", round(e["value"]), "pcm16", round(e.get("pcm16", {}).get("value", 0)), "h2d GB/s", round(e.get("h2d_GBps_measured", 0), 1))
print("  " + " ".join(f"{k['name']}={k['us']:.1f}" for k in d.get("kernels", [])))
