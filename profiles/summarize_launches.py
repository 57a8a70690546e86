"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of total)."""
import collections
import csv
import re
import sys


def summarize(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        k = re.sub(r"\(.*", "", row["Kernel Name"])[:90]
        v = float(row["Metric Value"].replace(",", ""))
        v = {"ns": v / 1e3, "us": v, "ms": v * 1e3}.get(row["Metric Unit"], v)
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    out = [f"| kernel | launches | total us | share |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% |")
    return "\n".join(out), tot


if __name__ == "__main__":
    table, tot = summarize(sys.argv[1])
    print(f"total {tot:.1f} us over the captured launches\n")
    print(table)
