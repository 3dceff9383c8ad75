"""`ncu --set full` reports (.ncu-rep) -> one JSON with the handful of metrics DESIGN.md quotes per captured launch.
Usage: python scripts/ncu_extract.py out.json rep1.ncu-rep [rep2.ncu-rep ...]   (needs `ncu` on PATH; reads reports, no GPU)"""
import csv
import io
import json
import subprocess
import sys

WANT = {
    "gpu__time_duration.sum": "time",
    "launch__registers_per_thread": "regs",
    "launch__shared_mem_per_block_dynamic": "smem_dyn",
    "launch__cluster_dim_x": "cluster_x",
    "dram__bytes_read.sum": "dram_read",
    "dram__bytes_write.sum": "dram_write",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed": "dram_pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active": "tensor_pipe_pct_active",
    "sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed": "tensor_pipe_pct_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed": "sm_pct",
    "lts__t_sector_hit_rate.pct": "l2_hit_pct",
    "lts__throughput.avg.pct_of_peak_sustained_elapsed": "l2_pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active": "occupancy_pct",
    "smsp__inst_executed.sum": "warp_insts",
    "sm__cycles_active.avg": "sm_cycles_active",
}


def extract(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    cols = {}
    for want, short in WANT.items():
        for i, h in enumerate(hdr):
            if h == want or h.endswith("." + want):
                cols[short] = i
                break
    out = []
    for r in data:
        e = {"report": path.split("/")[-1], "kernel": r[hdr.index("Kernel Name")].split("(")[0].replace("<unnamed>::", "").replace("void ", ""),
             "grid": r[hdr.index("Grid Size")], "block": r[hdr.index("Block Size")]}
        for short, i in cols.items():
            v = r[i].replace(",", "")
            try:
                e[short] = float(v)
            except ValueError:
                e[short] = v
            e[short + "_unit"] = units[i]
        out.append(e)
    return out


if __name__ == "__main__":
    allk = []
    for p in sys.argv[2:]:
        allk += extract(p)
    json.dump(allk, open(sys.argv[1], "w"), indent=1)
    for e in allk:
        print(e["report"], e["kernel"][:40], e["grid"], e["block"], {k: v for k, v in e.items() if k in ("time", "regs", "dram_read", "dram_write", "dram_pct", "tensor_pipe_pct_elapsed", "tensor_pipe_pct_active", "l2_hit_pct", "occupancy_pct")})
