#!/usr/bin/env python
"""Summarise an ncu report of the warp-specialised tcgen05 kernels: headline counters per launch and the source
lines (SASS) where warps stall most -- with their stall reasons -- so that "who waits for whom" can be read off.

    python profiles/ncu_roles.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "sm__cycles_elapsed.avg", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]


def main():
    rep = sys.argv[1]
    top_n = int(sys.argv[2]) if len(sys.argv) > 2 else 24
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        print("=" * 100)
        print(r[ix["Kernel Name"]][:110])
        for w in WANT:
            if w in ix:
                print(f"  {w:90s} {r[ix[w]]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    for part in src.split('"Kernel Name"')[1:]:
        rows = list(csv.reader(('"Kernel Name"' + part).splitlines()))
        hdr = rows[1]
        data = [r for r in rows[2:] if len(r) == len(hdr)]
        ix = {h: i for i, h in enumerate(hdr)}
        stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        tot = sum(int(r[ix["# Samples"]]) for r in data) or 1
        tinst = sum(int(r[ix["Instructions Executed"]]) for r in data)
        print("-" * 100)
        print(rows[0][1][:110], "| samples", tot, "| warp instructions", tinst)
        top = sorted(range(len(data)), key=lambda i: -int(data[i][ix["# Samples"]]))[:top_n]
        for i in sorted(top):
            r = data[i]
            st = {c[6:]: int(r[ix[c]]) for c in stall_cols if int(r[ix[c]]) > 0}
            st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
            print(f"  line {i:5d} {100 * int(r[ix['# Samples']]) / tot:5.1f}%  x{r[ix['Instructions Executed']]:>9s}  {r[ix['Source']].strip()[:60]:60s} {st}")


if __name__ == "__main__":
    main()
