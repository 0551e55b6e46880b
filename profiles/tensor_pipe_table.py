#!/usr/bin/env python
"""Tensor-pipe utilisation of every launch of ONE CQL update, from the ncu launch list written by profiles/r02_prof.sh
(gpu__time_duration.sum + sm__pipe_tensor_cycles_active + smsp__issue_active per launch).

    python profiles/tensor_pipe_table.py profiles/r02_launches.csv > profiles/r02_update_tensor_pipe.txt
"""
import collections
import csv
import sys

T = "gpu__time_duration.sum"
P = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
I = "smsp__issue_active.avg.pct_of_peak_sustained_active"


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    idx = {h: i for i, h in enumerate(rows[hi])}
    out = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(idx):
            continue
        d = out.setdefault(int(r[idx["ID"]]), {"name": r[idx["Kernel Name"]].split("(")[0].replace("void ", ""),
                                               "grid": r[idx["Grid Size"]]})
        v, u = float(r[idx["Metric Value"]].replace(",", "")), r[idx["Metric Unit"]]
        if r[idx["Metric Name"]] == T:
            v = v / 1000 if u in ("ns", "nsecond") else v * 1000 if u in ("ms", "msecond") else v
        d[r[idx["Metric Name"]]] = v
    return list(out.values())


def main(path):
    L = load(path)
    heads = [i for i, d in enumerate(L) if "k_step_head" in d["name"]]
    # a regular (graph-replayed) update: exactly one big critic forward -- cql_timed_update's steps launch it four times
    # and run bwd1 / bwd2 back to back on full grids, so they are not the product schedule
    segs = [L[a:b] for a, b in zip(heads[:-1], heads[1:])]
    regular = [g for g in segs if sum("tc_fwd_h2_kernel<3, 1>" in d["name"] for d in g) == 1]
    seg = (regular or segs)[-1]
    tot = sum(d[T] for d in seg)
    tc = [d for d in seg if d.get(P, 0) > 0.5]
    t_tc = sum(d[T] for d in tc)
    print(f"# one CQL update (batch 1024, precision f16x3) out of the ncu launch list {path}")
    print("# (bench.py --steps 2 --warmup 1 --no-cpu under ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active...,")
    print("#  smsp__issue_active... --clock-control none; per-launch times are cold-cache and serialised)")
    print(f"{'kernel':58s} {'grid':>14s} {'time':>9s} {'tensor pipe':>12s} {'issue slots':>12s}")
    for d in seg:
        print(f"{d['name'][:58]:58s} {d['grid']:>14s} {d[T]:7.1f} us {d.get(P, 0):10.2f} % {d.get(I, 0):10.1f} %")
    w = sum(d[T] * d[P] for d in tc) / t_tc
    big = sorted(tc, key=lambda d: -d[T])[:3]
    wb = sum(d[T] * d[P] for d in big) / sum(d[T] for d in big)
    print()
    print(f"{len(seg)} launches, {tot:.1f} us under ncu; {len(tc)} tcgen05 launches, {t_tc:.1f} us")
    print(f"tensor pipe, time-weighted over the tcgen05 launches: {w:.1f} % of active cycles")
    print("note: in the product schedule the critic's bwd1 (60 CTAs) and bwd2 (88 CTAs) run SIDE BY SIDE on disjoint SMs, and so do")
    print("      the actor's bwd1 / bwd2; ncu serialises them.  'tensor pipe' is % of the ACTIVE cycles of the SMs a launch occupies.")
    print(f"tensor pipe, time-weighted over the three big launches ({', '.join(d['name'].split('::')[-1][:16] for d in big)}): {wb:.1f} %")


if __name__ == "__main__":
    main(sys.argv[1])
