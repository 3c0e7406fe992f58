"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel: count, mean, share (dev aid)."""
import collections
import csv
import sys


def main():
    path = sys.argv[1]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    ci = {h: i for i, h in enumerate(rows[0])}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        if r[ci["Metric Name"]] != "gpu__time_duration.sum":
            continue
        name = r[ci["Kernel Name"]]
        name = name if name.startswith(("ppo::", "void ppo::")) else "torch: " + name.split("<")[0][-60:]
        v = float(r[ci["Metric Value"]].replace(",", ""))
        v *= {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ci["Metric Unit"]], 1.0)
        agg.setdefault(name[:100], []).append(v)
    tot = sum(sum(v) for v in agg.values())
    ours = sum(sum(v) for k, v in agg.items() if "ppo::" in k)
    print(f"# {path}: {sum(len(v) for v in agg.values())} launches, {tot:.1f} us total, {ours / tot * 100:.1f}% in ppo:: kernels")
    print("# (ncu serialises launches and runs them cold: compare shares, not absolutes)")
    print(f"{'count':>5} {'mean us':>10} {'share':>7} {'share of ppo::':>15}  kernel")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        so = f"{sum(v) / ours * 100:14.1f}%" if "ppo::" in k else " " * 15
        print(f"{len(v):5d} {sum(v) / len(v):10.1f} {sum(v) / tot * 100:6.1f}% {so}  {k}")


if __name__ == "__main__":
    main()
