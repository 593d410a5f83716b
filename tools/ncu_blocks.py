"""Summarises an .ncu-rep: key metrics + basic-block breakdown of executed instructions (SASS view).
    python tools/ncu_blocks.py gpurun_out/prof.ncu-rep [top_n]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    keys = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
            "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
            "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
            "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
            "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
            "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct"]
    for h, u, v in zip(hdr, units, vals):
        if h in keys:
            print(f"{h} [{u}] = {v}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))[2:]
    total = sum(int(r[5]) for r in rows)
    print(f"total warp instructions {total / 1e9:.3f} G over {len(rows)} SASS instructions")
    blocks, cur = [], None
    for i, r in enumerate(rows):
        n = int(r[5])
        if cur and cur["n"] == n:
            cur["end"] = i
            cur["cnt"] += 1
            cur["samples"] += int(r[4])
        else:
            cur = {"start": i, "end": i, "n": n, "cnt": 1, "samples": int(r[4])}
            blocks.append(cur)
    blocks.sort(key=lambda b: -b["n"] * b["cnt"])
    for b in blocks[:top]:
        ops = []
        for j in range(b["start"], b["end"] + 1):
            t = rows[j][1].split()
            ops.append((t[1] if t[0].startswith("@") else t[0]).split(".")[0])
        c = collections.Counter(ops)
        print(f"{b['start']:5d}-{b['end']:5d} exec={b['n'] / 1e6:8.2f}M x{b['cnt']:3d} = {b['n'] * b['cnt'] / total * 100:5.1f}% "
              f"samples {b['samples']:6d} {dict(c.most_common(7))}")


if __name__ == "__main__":
    main()
