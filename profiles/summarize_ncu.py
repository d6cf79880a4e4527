"""Text summary of an `ncu --set full` report: per-kernel headline metrics + SASS opcode mix
(`ncu -i rep --page raw --csv`, `--page source --csv`).  Usage: summarize_ncu.py rep.ncu-rep [units_per_launch]"""
import collections
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "launch__waves_per_multiprocessor",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("== " + r[hdr.index("Kernel Name")])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"  {k:82s} {r[i]:>18s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = src.split('"Kernel Name"')
    for b in blocks[1:]:
        lines = ('"Kernel Name"' + b).splitlines()
        name = lines[0]
        rr = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
        h = rr[0]
        if "Instructions Executed" not in h:
            continue
        ie, isrc = h.index("Instructions Executed"), h.index("Source")
        ist = h.index("Warp Stall Sampling (All Samples)")
        cnt, st = collections.Counter(), collections.Counter()
        for r in rr[1:]:
            if len(r) <= ie or not r[ie].isdigit():
                continue
            s = r[isrc].strip()
            op = (s.split()[1] if s.startswith("@") else s.split()[0]).split(".")[0]
            cnt[op] += int(r[ie])
            st[op] += int(r[ist]) if r[ist].isdigit() else 0
        tot, tots = sum(cnt.values()), max(sum(st.values()), 1)
        print("-- SASS opcode mix (warp-level instructions executed) " + name[:120])
        for op, c in cnt.most_common(14):
            print(f"  {op:10s} {c:14d} {100 * c / tot:5.1f}%   stall samples {100 * st[op] / tots:5.1f}%")
        print(f"  {'TOTAL':10s} {tot:14d}")


if __name__ == "__main__":
    main()
