"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel share of the captured device time."""
import collections
import csv
import sys


def main(path, title):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot, n = 0.0, 0
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        ms = v / 1e6 if unit in ("ns", "nsecond") else (v / 1e3 if unit.startswith("u") else v)
        agg[row["Kernel Name"]][0] += 1
        agg[row["Kernel Name"]][1] += ms
        tot += ms
        n += 1
    ours = {k: v for k, v in agg.items() if "msf::" in k}
    print(f"# {title}\n")
    print(f"`ncu --metrics gpu__time_duration.sum --clock-control none` (per-launch times are cold-cache and serialised: compare SHARES).")
    print(f"Captured {n} launches, {tot:.1f} ms of kernel time; kernels of this repo: {sum(v[0] for v in ours.values())} launches, "
          f"**{100 * sum(v[1] for v in ours.values()) / tot:.1f} %** of the time.\n")
    print("## Kernels of this repo\n\n| share | launches | avg us | kernel |\n|---|---|---|---|")
    for k, v in sorted(ours.items(), key=lambda kv: -kv[1][1]):
        print(f"| {100 * v[1] / tot:.2f}% | {v[0]} | {1e3 * v[1] / v[0]:.1f} | `{k[:110]}` |")
    print("\n## Top 25 overall\n\n| share | launches | avg us | kernel |\n|---|---|---|---|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:25]:
        print(f"| {100 * v[1] / tot:.2f}% | {v[0]} | {1e3 * v[1] / v[0]:.1f} | `{k[:110]}` |")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else sys.argv[1])
