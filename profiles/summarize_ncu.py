#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU needed) into profiles/<name>.json + a raw-page CSV of the chosen metrics.

usage: python profiles/summarize_ncu.py gpurun_out/prof.ncu-rep profiles/r1_step_bf16_lut1 [traffic_key]
If traffic_key is given, profiles/traffic.json[traffic_key] = dram bytes (read+write) per launch, which bench.py
reports as roofline.traffic.
"""
import csv
import io
import json
import os
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct",
    "l1tex__t_bytes_pipe_lsu_mem_global_op_ld.sum", "l1tex__t_bytes_pipe_lsu_mem_global_op_st.sum",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
    "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
    "launch__waves_per_multiprocessor", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "smsp__issue_active.avg.pct_of_peak_sustained_active",
]
UNIT_SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12, "ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}


def main():
    rep, out = sys.argv[1], sys.argv[2]
    key = sys.argv[3] if len(sys.argv) > 3 else None
    raw = subprocess.check_output(["ncu", "-i", rep, "--page", "raw", "--csv"], text=True)
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    launches = []
    for r in data:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k in KEEP:
            if k in hdr:
                i = hdr.index(k)
                try:
                    v = float(r[i].replace(",", ""))
                except ValueError:
                    continue
                d[k] = v * UNIT_SCALE.get(units[i], 1)
        d["dram_bytes_total"] = d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0)
        if d.get("gpu__time_duration.sum"):
            d["dram_GBps"] = d["dram_bytes_total"] / d["gpu__time_duration.sum"] / 1e9
        launches.append(d)
    json.dump({"source": os.path.basename(rep), "note": "per-launch values under ncu (serialised, cold cache)", "launches": launches},
              open(out + ".json", "w"), indent=1)
    with open(out + ".raw.csv", "w") as f:
        w = csv.writer(f)
        idx = [hdr.index(k) for k in ["Kernel Name"] + [k for k in KEEP if k in hdr]]
        for r in rows:
            w.writerow([r[i] for i in idx])
    if key:
        tp = os.path.join(os.path.dirname(os.path.abspath(__file__)), "traffic.json")
        t = json.load(open(tp)) if os.path.exists(tp) else {}
        t[key] = {"bytes": sum(l["dram_bytes_total"] for l in launches) / len(launches), "source": "profiles/" + os.path.basename(out) + ".json",
                  "kernel": launches[0]["kernel"][:80], "launches_averaged": len(launches)}
        json.dump(t, open(tp, "w"), indent=1)
    for l in launches:
        print(l["kernel"][:60], "%.1f us" % (l["gpu__time_duration.sum"] * 1e6), "%.3f GB dram" % (l["dram_bytes_total"] / 1e9), "%.0f GB/s" % l.get("dram_GBps", 0))


if __name__ == "__main__":
    main()
