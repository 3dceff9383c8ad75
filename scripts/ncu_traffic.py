"""ncu CSV (dram__bytes_read.sum, dram__bytes_write.sum, gpu__time_duration.sum per launch) -> profiles/<round>_tc_traffic.json.
bench.py reads that file to fill roofline.traffic (DRAM bytes per launch of the dominant kernel, same averaging as `achieved`).
Usage: python scripts/ncu_traffic.py gpurun_out/tc_traffic.csv profiles/r01_tc_traffic.json [kernel-substring]"""
import collections
import csv
import json
import sys


def main(src, dst, needle="tc_gemm_kernel"):
    rows = list(csv.reader(open(src)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, ui, vi, ii = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Unit"), hdr.index("Metric Value"), hdr.index("ID")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6, "nsecond": 1e-3, "usecond": 1.0, "msecond": 1e3, "second": 1e6}
    per = collections.defaultdict(dict)
    for r in rows[hi + 1:]:
        if len(r) <= vi or not any(nd in r[ki] for nd in needle.split(",")):
            continue
        per[r[ii]][r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    n = len(per)
    rd = sum(v.get("dram__bytes_read.sum", 0.0) for v in per.values())
    wr = sum(v.get("dram__bytes_write.sum", 0.0) for v in per.values())
    us = sum(v.get("gpu__time_duration.sum", 0.0) for v in per.values())
    out = {"kernel": needle, "source": src, "launches": n, "dram_read_bytes_per_step": rd, "dram_write_bytes_per_step": wr,
           "dram_bytes_per_launch": (rd + wr) / n if n else None, "ncu_time_us_per_step": us,
           "note": "one eager SD1.5 UNet step under ncu (cold caches, serialised launches): shares and byte counts are meaningful, absolute times are not"}
    json.dump(out, open(dst, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main(*sys.argv[1:])
