#!/usr/bin/env python
"""Summarise an ncu --set full report (.ncu-rep) into the small tables kept under profiles/.

  python tools/ncu_summary.py gpurun_out/r01_v3_legacy.ncu-rep profiles/r01_v3_ncu_full_tcresnet8_b512

writes <out>.csv (one row per profiled launch: duration, DRAM bytes, registers, grid, issue/FMA pipe utilisation and the
three largest warp-stall reasons) and <out>_traffic.json (kernel -> mean DRAM bytes per launch; bench.py reads it for
`roofline.traffic`).  Numbers in these files are taken UNDER the profiler (cold caches, serialised launches): they are
evidence for traffic, occupancy and stall reasons, never bench values.
"""
import collections
import csv
import io
import json
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"^void ", "", name)
    return re.sub(r"\(.*$", "", name)


def main():
    rep, out = sys.argv[1], sys.argv[2]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, body = rows[0], rows[1], rows[2:]
    ix = {n: i for i, n in enumerate(hdr)}
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}

    def val(row, key):
        return float(row[ix[key]].replace(",", "")) * scale.get(units[ix[key]], 1.0)

    stalls = [k for k in hdr if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("per_issue_active.ratio")
              and "not_issued" not in k]
    table, traffic = [], collections.defaultdict(list)
    for r in body:
        name = short(r[ix["Kernel Name"]])
        rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
        top = sorted(((float(r[ix[k]]), k.split("stalled_")[1].replace("_per_issue_active.ratio", "")) for k in stalls), reverse=True)
        top = [t for t in top if t[1] != "selected"][:3]
        table.append([name, f"{val(r, 'gpu__time_duration.sum'):.2f}", int(rd), int(wr), r[ix["launch__grid_size"]],
                      r[ix["launch__block_size"]], r[ix["launch__registers_per_thread"]],
                      f"{float(r[ix['sm__warps_active.avg.pct_of_peak_sustained_active']]):.1f}",
                      f"{float(r[ix['smsp__issue_active.avg.pct_of_peak_sustained_active']]):.1f}",
                      f"{float(r[ix['sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active']]):.1f}",
                      int(float(r[ix["smsp__inst_executed.sum"]])), " ".join(f"{n}={v:.2f}" for v, n in top)])
        traffic[name].append(rd + wr)
    with open(out + ".csv", "w", newline="") as f:
        w = csv.writer(f)
        w.writerow(["kernel", "duration_us_under_ncu", "dram_read_bytes", "dram_write_bytes", "grid", "block", "regs", "warps_active_pct",
                    "issue_active_pct", "fma_pipe_active_pct", "warp_instructions", "top_stalls_per_issue"])
        w.writerows(table)
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from bench import sources_sha          # hash of the kernel sources this capture describes; bench.py refuses a stale one
    json.dump({"source": rep.split("/")[-1], "kernels_sha": sources_sha(),
               "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu --set full)",
               "bytes_per_launch": {k: sum(v) / len(v) for k, v in traffic.items()}}, open(out + "_traffic.json", "w"), indent=1)
    print(f"{len(table)} launches -> {out}.csv")


if __name__ == "__main__":
    main()
