#!/usr/bin/env python
"""Turns `ncu --set full` captures of k_render into profiles/r02_ncu_k_render.json — the file bench.py reads the
per-launch DRAM traffic and issue-slot utilisation from (so that those figures in the bench line are never constants in
bench.py). Records the hash of the sources the capture was taken with (bench.source_sha16: the capture visit runs
`python -c "import bench; print(bench.source_sha16())"` on the box, next to the .ncu-rep).

    python tools/ncu_to_json.py <source_sha16> workload=path.ncu-rep [workload=path.ncu-rep ...]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {
    "dram__bytes_read.sum": "dram_read_bytes", "dram__bytes_write.sum": "dram_write_bytes", "gpu__time_duration.sum": "ncu_duration",
    "smsp__issue_active.avg.pct_of_peak_sustained_active": "issue_active_pct", "smsp__inst_executed.sum": "warp_instructions",
    "smsp__thread_inst_executed_per_inst_executed.ratio": "lanes_active", "sm__warps_active.avg.pct_of_peak_sustained_active": "warps_active_pct",
    "launch__registers_per_thread": "registers", "sass__inst_executed_local_loads": "local_loads", "sass__inst_executed_local_stores": "local_stores",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio": "stall_long_scoreboard",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct", "l1tex__t_sector_hit_rate.pct": "l1_hit_pct",
}
SCALE = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1.0}


def read(rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    out = {"kernel": vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else None}
    for h, u, v in zip(hdr, units, vals):
        if h in KEYS:
            x = float(v.replace(",", ""))
            if h.startswith("dram__bytes"):
                x *= SCALE.get(u, 1.0)
            out[KEYS[h]] = x
            if h == "gpu__time_duration.sum":
                out["ncu_duration_unit"] = u
    out["dram_bytes"] = out.get("dram_read_bytes", 0.0) + out.get("dram_write_bytes", 0.0)
    out["issue_active"] = out.get("issue_active_pct", 0.0) / 100.0
    return out


def main():
    sha = sys.argv[1]
    doc = {"source_sha16": sha, "how": "ncu --set full --clock-control none --import-source on -k regex:k_render --launch-skip 1 -c 1 "
                                      "python tools/profile_frame.py <workload> 2 (tools/gpu_*.sh), read here with ncu -i --page raw --csv",
           "captures": {}}
    for arg in sys.argv[2:]:
        name, rep = arg.split("=", 1)
        rec = read(rep)
        rec["report"] = os.path.basename(rep)
        doc["captures"][name] = rec
    path = os.path.join(ROOT, "profiles", "r02_ncu_k_render.json")
    with open(path, "w") as fh:
        json.dump(doc, fh, indent=1)
    print(f"wrote {path}: {list(doc['captures'])}")


if __name__ == "__main__":
    main()
