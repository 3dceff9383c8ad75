"""ncu launch list (gpu__time_duration.sum CSV) -> per-kernel summary table (markdown) for profiles/."""
import collections
import csv
import re
import sys


def summarize(path):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr, data = rows[hi], rows[hi + 1:]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    mi = hdr.index("Metric Name") if "Metric Name" in hdr else None
    agg = collections.defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for r in data:
        if len(r) <= vi or (mi is not None and r[mi] != "gpu__time_duration.sum"):
            continue
        m = re.search(r"(\w+_kernel)\b", r[ki])
        key = m.group(1) if m else r[ki][:40]
        v = float(r[vi].replace(",", "")) / 1000.0
        agg[key][0] += 1
        agg[key][1] += v
        tot += v
    out = [f"| kernel | launches | total us | avg us | share |", "|---|---:|---:|---:|---:|"]
    for k, (n, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
        out.append(f"| `{k}` | {n} | {t:.1f} | {t / n:.1f} | {100 * t / tot:.1f}% |")
    out.append(f"| **total** | {sum(a[0] for a in agg.values())} | {tot:.1f} | | |")
    return "\n".join(out)


if __name__ == "__main__":
    print(summarize(sys.argv[1]))
