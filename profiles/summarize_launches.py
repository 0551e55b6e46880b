#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list of `bench.py --steps 2 --warmup 1 --no-cpu`:
shares of one CQL update (between two k_step_head launches) and of the whole capture.

    python profiles/summarize_launches.py profiles/r01_f16x3_launches.csv > profiles/r01_f16x3_launch_summary.txt
"""
import collections
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    idx = {h: i for i, h in enumerate(hdr)}
    out = []
    for r in rows[hi + 1:]:
        if len(r) < len(hdr) or r[idx["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[idx["Kernel Name"]].split("(")[0][:70]
        v, u = float(r[idx["Metric Value"]]), r[idx["Metric Unit"]]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        out.append((name, v))
    return out


def table(seg):
    agg = collections.OrderedDict()
    for n, v in seg:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(a[1] for a in agg.values())
    return tot, sorted(agg.items(), key=lambda kv: -kv[1][1])


def main(path):
    launches = load(path)
    heads = [i for i, (n, _) in enumerate(launches) if "k_step_head" in n]
    # a regular (graph-replayed) update: exactly one big critic forward -- cql_timed_update's steps launch it four times
    # and run bwd1 / bwd2 back to back on full grids, so they are not the product schedule
    segs = [launches[a:b] for a, b in zip(heads[:-1], heads[1:])]
    regular = [g for g in segs if sum("tc_fwd_h2_kernel<3, 1>" in n for n, _ in g) == 1]
    seg = (regular or segs)[-1]
    tot, rows = table(seg)
    print(f"# ncu --metrics gpu__time_duration.sum --clock-control none  python bench.py --steps 2 --warmup 1 --no-cpu  (precision f16x3, default)")
    print(f"# raw list: {path}.  Per-launch times are cold-cache and serialised: compare SHARES, not absolutes.")
    print(f"# one CQL update (batch 1024, ML-20M table): {len(seg)} launches, {tot:.1f} us under ncu")
    for k, a in rows:
        print(f"{a[1]:8.1f} us {a[0]:3d}x {100 * a[1] / tot:5.1f}%  {k}")
    tot, rows = table(launches)
    print()
    print("# whole capture (MDP build + warm-up + timed updates + timed_update x5 + e2e + scoring + top-k filter + sampler sweep), by kernel")
    for k, a in rows[:30]:
        print(f"{a[1]:10.1f} us {a[0]:4d}x {100 * a[1] / tot:5.1f}%  {k}")


if __name__ == "__main__":
    main(sys.argv[1])
