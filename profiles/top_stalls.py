"""Summarise `ncu -i X.ncu-rep --page source --csv` output: instructions with the most stall samples."""
import csv
import sys


def load(path):
    rows = list(csv.reader(open(path)))
    # the file may hold several kernels: header rows start with "Address"
    out, hdr = [], None
    for r in rows:
        if r and r[0] == "Address":
            hdr = r
            continue
        if hdr is None or len(r) < len(hdr) or not r[0].startswith("0x"):
            continue
        out.append(dict(zip(hdr, r)))
    return out


if __name__ == "__main__":
    data = load(sys.argv[1])
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
    tot = sum(int(d["# Samples"] or 0) for d in data)
    print("total samples", tot, "instructions", sum(int(d["Instructions Executed"] or 0) for d in data))
    stall_cols = [k for k in data[0] if k.startswith("stall_") and "Not Issued" not in k]
    agg = {k: sum(int(d[k] or 0) for d in data) for k in stall_cols}
    print("stall mix:", {k: v for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
    for i, d in sorted(enumerate(data), key=lambda t: -int(t[1]["# Samples"] or 0))[:n]:
        s = int(d["# Samples"] or 0)
        top = max(stall_cols, key=lambda k: int(d[k] or 0))
        print(f"{s:6d} {100 * s / max(tot, 1):5.1f}%  ex={int(d['Instructions Executed'] or 0):7d} #{i:5d} {top:16s} {d['Source'].strip()[:80]}")
