#!/usr/bin/env python
"""bench.py -- SD 1.5 UNet 512x512 bs=1 denoise-step throughput on N B200s (BASELINE.json config[1] hot path).

One "step" = one UNet Model::run() on one 4x64x64 latent (the reference runs two per sampler step with CFG; the
metric counts UNet runs).  fp16 weights / fp16 arithmetic / attention fusion, i.e. what sd.cpp sets
(src/sd.cpp:1616-1683).  Synthetic SD1.5-shaped graph + seeded random weights (no checkpoints offline).

  value  : steps/s with weights and inputs resident in HBM (one captured CUDA graph per step), CUDA-event timed
  e2e    : the same metric through the reference-facing C ABI with HOST buffers: every step pushes the inputs from
           pinned host memory, streams all 1.72 GB of weights pinned-host -> HBM ring (the CUDA WeightsProvider; N>1:
           rank 0 uploads, NCCL broadcast) and reads the output latent back
  roofline: the dominant kernel (tcgen05 implicit-GEMM conv / GEMM) timed per launch with CUDA events in an eager pass
  cpu_baseline: the reference's own CPU path (oracle/_ref: reference sources + XNNPACK) on the host cores

`--impl reference` times the reference CPU implementation alone (bounded sample per step).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RANK = int(os.environ.get("RANK", "0"))
LOCAL_RANK = int(os.environ.get("LOCAL_RANK", "0"))
WORLD = int(os.environ.get("WORLD_SIZE", "1"))
if WORLD > 1 and "OSB_KEEP_VISIBLE" not in os.environ:
    # one process per GPU: each rank sees exactly its own device as device 0 (engine and torch alike)
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    devs = vis.split(",") if vis else [str(i) for i in range(64)]
    os.environ["CUDA_VISIBLE_DEVICES"] = devs[LOCAL_RANK]

# The contract is ONE JSON line on stdout: native libraries (NCCL prints its version banner) must not pollute it, so
# fd 1 points at stderr while the bench runs and the line is written to the saved descriptor at the end.
_REAL_STDOUT = os.dup(1)
os.dup2(2, 1)


def emit_line(obj):
    os.write(_REAL_STDOUT, (json.dumps(obj) + "\n").encode())


import numpy as np  # noqa: E402

from onnxstream_b200 import emit  # noqa: E402
from onnxstream_b200.model import Model, ENGINE_LIB  # noqa: E402

ORACLE_LIB = os.path.join(ROOT, "oracle", "_ref", "liboracle_ref.so")
REF_THREADS = min(os.cpu_count() or 1, 32)
os.environ.setdefault("OMP_NUM_THREADS", str(REF_THREADS))
METRIC = "SD1.5 UNet 512x512 bs=1 denoise steps/s (one UNet Model::run per step)"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d.get("bf16_tflops_sustained", 1450.4), d.get("hbm_gbs", 6567.7), "measured (MEASURED_PEAKS.json, sustained bf16 cuBLAS)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


def model_dir(workload: str) -> str:
    base = "/dev/shm" if os.path.isdir("/dev/shm") and os.access("/dev/shm", os.W_OK) else "/tmp"
    return os.path.join(base, f"osb200_bench_{workload}") + "/"


def ensure_model(workload: str):
    """Emit the synthetic model directory once per box (rank 0), return (dir, cfg, flops, weight_bytes)."""
    cfg = emit.UNetConfig.sd15(64) if workload == "sd15_unet_fp16" else emit.UNetConfig.tiny(16)
    d = model_dir(workload)
    meta = os.path.join(d, "meta.json")
    if not os.path.exists(meta):
        if RANK == 0:
            g = emit.emit_unet(d, cfg, "float16", seed=0)
            json.dump({"flops": g.flops, "weight_bytes": g.weight_bytes, "params": g.weight_params, "ops": len(g.lines)}, open(meta + ".tmp", "w"))
            os.replace(meta + ".tmp", meta)
        else:
            while not os.path.exists(meta):
                time.sleep(0.5)
    m = json.load(open(meta))
    return d, cfg, m


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self):
        self.rows = []
        self.proc = None

    def start(self):
        try:
            q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
            self.proc = subprocess.Popen(["nvidia-smi", "-i", os.environ.get("CUDA_VISIBLE_DEVICES", "0").split(",")[0], f"--query-gpu={q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 7 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 7 for i in range(4) if r[3 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm)}


def dist_setup():
    if WORLD == 1:
        return None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", device_id=torch.device("cuda", 0))
    return dist


def dist_max(dist, x: float) -> float:
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def dist_barrier(dist):
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


def make_engine_model(d, wp, resident, graph, comm=None):
    m = Model(ENGINE_LIB, 0, wp)
    m.set_option("use_fp16_arithmetic", True)
    m.set_option("fuse_ops_in_attention", True)
    m.lib.model_set_option(m.h, b"b200_resident_weights", 1 if resident else 0)
    m.lib.model_set_option(m.h, b"b200_cuda_graph", 1 if graph else 0)
    if comm is not None:
        m.lib.model_b200_set_comm(m.h, comm, RANK, WORLD)
    m.read_file(d + "model.txt")
    return m


def step_api(m, inputs, out_name="out_5F_sample"):
    m.clear_tensors()
    for k, v in inputs.items():
        m.add_tensor(k, v)
    m.run()
    return m.get_tensor(out_name)


def model_lines_flops(d):
    """Per-line FLOPs of Conv / MatMul / Gemm ops of a model.txt (for bounded CPU samples)."""
    import re
    out = []
    for line in open(d + "model.txt").read().splitlines():
        if not line:
            continue
        typ = line.split("*", 1)[0].split(":")[1]
        fl = 0
        if typ in ("Conv", "MatMul", "Gemm"):
            ins = line.split("*input:")[1].split("*output:")[0].split(";")
            outs = line.split("*output:")[1].split("*")[0].split(";")

            def shape(t):
                s = t[t.index("(") + 1:-1]
                s = s.split(":")[-1]
                return [int(x) for x in s.split(",") if x]
            o = shape(outs[0])
            if typ == "Conv":
                w = shape(ins[1])
                fl = 2 * int(np.prod(o)) * w[1] * w[2] * w[3]
            else:
                a = shape(ins[0])
                fl = 2 * int(np.prod(o)) * a[-1]
        out.append((line, fl))
    return out


def cpu_sample(d, inputs, frac, tag):
    """Time the reference CPU path on a prefix of the graph holding ~frac of its FLOPs; returns (seconds, actual fraction)."""
    lf = model_lines_flops(d)
    total = sum(f for _, f in lf)
    acc, n = 0, 0
    for i, (_, f) in enumerate(lf):
        acc += f
        n = i + 1
        if acc >= frac * total:
            break
    if frac >= 0.999:
        n, acc = len(lf), total
    fn = d + f"model_prefix_{tag}.txt"
    open(fn, "w").write("\n".join(l for l, _ in lf[:n]) + "\n")
    # XNNPACK's pthreadpool with one worker per core is far slower than a moderate pool on many-core hosts (measured: 128
    # workers -> 295 s per UNet run on the GPU box, vs 38 s with 8 on an 8-core host), so the reference gets min(cores, 32).
    m = Model(ORACLE_LIB, REF_THREADS, "nocache")
    m.set_option("fuse_ops_in_attention", True)   # fp16 weights, fp32 arithmetic: the reference's --rpi mode (src/sd.cpp:1636-1638)
    m.read_file(fn)
    return m, acc / total, n


def run_reference(args):
    d, cfg, meta = ensure_model(args.workload)
    inputs = emit.unet_inputs(cfg)
    cores = REF_THREADS
    if RANK != 0:
        return
    if not os.path.exists(ORACLE_LIB):
        emit_line({"impl": "reference", "unavailable": "oracle/_ref/liboracle_ref.so missing (built from /root/reference by __graft_entry__.build())"})
        return
    # calibrate on a small prefix, then size the per-step sample so that (K+W) steps take ~150 s
    m, f0, _ = cpu_sample(d, inputs, 0.03, "cal")
    t = time.time(); step_api(m, inputs, "nonexistent"); t_cal = time.time() - t
    t = time.time(); step_api(m, inputs, "nonexistent"); t_cal = min(t_cal, time.time() - t)
    full_est = t_cal / f0
    budget = 150.0 / (args.steps + args.warmup)
    frac = min(1.0, max(0.02, budget / full_est))
    m, f, nops = cpu_sample(d, inputs, frac, "ref")
    for _ in range(args.warmup):
        step_api(m, inputs, "nonexistent")
    t = time.time()
    for _ in range(args.steps):
        step_api(m, inputs, "nonexistent")
    dt = (time.time() - t) / args.steps
    value = f / dt
    sample = f"first {nops} of {meta['ops']} graph ops = {100 * f:.1f}% of the UNet's Conv/MatMul FLOPs per step, extrapolated to one full UNet run"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "SD1.5 UNet-shaped graph, 4x64x64 latent, 77x768 context, fp16 weights / fp32 XNNPACK arithmetic on host cores"},
        "cpu_baseline": {"value": value, "unit": "steps/s", "cores": cores, "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit_line(line)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="sd15_unet_fp16", choices=["sd15_unet_fp16", "tiny_unet_fp16"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    if args.impl == "reference":
        run_reference(args)
        return

    dist = dist_setup()
    d, cfg, meta = ensure_model(args.workload)
    dist_barrier(dist)
    inputs = emit.unet_inputs(cfg, seed=RANK)   # rank r = sample r (seed + rank, src/sd.cpp:2671)
    peak_tf, peak_bw, peak_src = peaks()

    # ---------------- value arm: weights + inputs resident in HBM, one CUDA graph per step ----------------
    mv = make_engine_model(d, "ram+nocache", resident=True, graph=True)
    for _ in range(3):      # run 1 fills the HBM weight cache, run 2 warms scratch, run 3 captures the graph
        out_v = step_api(mv, inputs)
    st_v = mv.stats()
    launches_per_step = int(st_v["kernel_launches"])
    clocks = ClockSampler(); clocks.start()      # sampled from the warm-up through both timed regions (value and e2e)
    mv.run_resident(args.warmup)
    dist_barrier(dist)
    gpu_ms = mv.run_resident(args.steps)
    dist_barrier(dist)
    gpu_ms = dist_max(dist, gpu_ms)
    ms_per_step = gpu_ms / args.steps
    value = WORLD * 1000.0 / ms_per_step

    # ---------------- supplementary: the same C-ABI call with host buffers but HBM-resident weights (the reference's --ram mode) ----
    # inputs H2D from pinned memory and output D2H every step, CUDA-graph replay in between; reported beside the streaming e2e
    e2e_resident = None
    e2e_r_s, e2e_r_err = float("inf"), None
    try:                      # no collectives inside the try: a rank-local failure must not strand the other ranks at a barrier
        for _ in range(2):
            step_api(mv, inputs)
        t0r = time.perf_counter()
        for _ in range(args.steps):
            step_api(mv, inputs)
        e2e_r_s = time.perf_counter() - t0r
    except Exception as e:    # supplementary only: never take the bench down
        e2e_r_err = str(e)
    e2e_r_s = dist_max(dist, e2e_r_s)          # every rank takes part, whatever happened above
    if e2e_r_err is None and e2e_r_s != float("inf"):
        e2e_resident = {"value": WORLD * args.steps / e2e_r_s, "unit": "steps/s", "ms_per_step": 1000.0 * e2e_r_s / args.steps,
                        "note": "host inputs / outputs through the C ABI every step, weights resident in HBM (b200_resident_weights, the reference's --ram mode); ranks not barrier-aligned"}
    else:
        e2e_resident = {"value": None, "note": f"failed: {e2e_r_err}"}

    # ---------------- roofline leg: eager pass with per-launch CUDA events on the tcgen05 kernel ----------------
    mv.lib.model_set_option(mv.h, b"b200_cuda_graph", 0)
    step_api(mv, inputs)
    mv.lib.osb_tc_profile(1)
    step_api(mv, inputs)
    prof = (ctypes.c_double * 4)()
    rc = mv.lib.osb_tc_profile_read(prof)
    if os.environ.get("OSB_TC_DUMP"):
        buf = ctypes.create_string_buffer(1 << 20)
        mv.lib.osb_tc_profile_dump.argtypes = [ctypes.c_char_p, ctypes.c_int]
        nb = mv.lib.osb_tc_profile_dump(buf, len(buf))
        open(os.environ["OSB_TC_DUMP"], "w").write(buf.raw[:max(nb, 0)].decode())
    mv.lib.osb_tc_profile(0)
    st_e = mv.stats()
    n_tc, tc_ms, tc_flops, tc_bytes = [float(x) for x in prof]
    achieved_tf = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    roofline = {"bound": "tensor", "kernel": "tc_gemm_kernel (tcgen05 implicit-GEMM conv / GEMM, all launches of one step)",
                "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf if peak_tf else None,
                "peak_source": peak_src, "launches": int(n_tc), "kernel_ms_per_step": tc_ms, "flops_per_step": tc_flops,
                "algorithmic_bytes_per_step": tc_bytes,
                # share of the SAME eager, event-instrumented step (numerator and denominator carry the same per-launch event cost);
                # compare with the kernel's share in profiles/r01_launches_step_final.csv
                "share_of_step": tc_ms / float(st_e["last_gpu_ms"]) if st_e.get("last_gpu_ms") else None, "traffic": None}
    # DRAM traffic of the same kernel from the committed ncu pass (profiles/r01_tc_traffic.json, made by scripts/ncu_traffic.py):
    # per launch, like `achieved`; the algorithmic bytes per launch stand beside it
    tpath = os.path.join(ROOT, "profiles", "r01_tc_traffic.json")
    if os.path.exists(tpath) and n_tc:
        tj = json.load(open(tpath))
        roofline["traffic"] = tj.get("dram_bytes_per_launch")
        roofline["traffic_unit"] = "bytes/launch (ncu dram__bytes_read.sum + dram__bytes_write.sum, %d launches of one step)" % tj.get("launches", 0)
        roofline["algorithmic_bytes_per_launch"] = tc_bytes / n_tc
    del mv

    # ---------------- e2e arm: host buffers through the C ABI, weights streamed every step ----------------
    comm = None
    me = Model(ENGINE_LIB, 0, "ram")
    if WORLD > 1:
        import torch.distributed as tdist
        ident = ctypes.create_string_buffer(128)
        if RANK == 0:
            assert me.lib.osb_comm_unique_id(ident) == 0
        obj = [bytes(ident.raw)]
        tdist.broadcast_object_list(obj, src=0)
        comm = me.lib.osb_comm_init(WORLD, RANK, obj[0])
        assert comm, "ncclCommInitRank failed"
    me.close()
    me = make_engine_model(d, "ram+nocache", resident=False, graph=False, comm=comm)
    for _ in range(max(2, args.warmup)):
        out_e = step_api(me, inputs)
    dist_barrier(dist)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        out_e = step_api(me, inputs)
    dist_barrier(dist)
    e2e_s = dist_max(dist, time.perf_counter() - t0)
    clk = clocks.stop()
    st = me.stats()
    e2e_value = WORLD * args.steps / e2e_s
    h2d = int(st["weight_bytes_streamed"] + st["h2d_input_bytes"])
    d2h = int(st["d2h_output_bytes"])
    parity = float(np.abs(out_e - out_v).max())

    line = {
        "metric": METRIC, "value": value, "unit": "steps/s", "n_gpus": WORLD, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f16", "data": "synthetic",
        "config": {"workload": "SD1.5 UNet-shaped graph (BASELINE config[1] hot path): %d ops, %.0f M params, %.3f TFLOP/step, 4x64x64 latent, 77x768 context, fp16 weights + fp16 arithmetic (fp32 accumulate), one sample per GPU" % (meta["ops"], meta["params"] / 1e6, meta["flops"] / 1e12),
                   "weights": "value: HBM-resident + CUDA graph; e2e: streamed pinned-host -> HBM ring every step",
                   "l2": "per-step working set = 1.72 GB of weights >> 126 MB L2, no flush needed", "parallelism": f"dp{WORLD} (rank r = sample r)"},
        "clocks": clk,
        "e2e": {"value": e2e_value, "unit": "steps/s", "ms_per_step": 1000.0 * e2e_s / args.steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                "weight_ring_bytes": int(st["weight_ring_bytes"]), "largest_node_bytes": int(st["weight_largest_node_bytes"]),
                "peak_hbm_resident_weight_bytes": int(st["weight_ring_bytes"]), "h2d_gbs": h2d / (e2e_s / args.steps) / 1e9,
                "max_abs_diff_vs_resident_arm": parity},
        "e2e_resident_weights": e2e_resident,
        "gpu_launches": launches_per_step * args.steps,
        "gpu_launches_per_step": launches_per_step, "tcgen05_launches_per_step": int(st_e["tc_launches"]),
        "roofline": roofline,
    }

    if RANK == 0 and WORLD == 1 and not args.no_cpu_baseline and os.path.exists(ORACLE_LIB):
        try:
            frac = 0.15
            m, f, nops = cpu_sample(d, inputs, frac, "cpu")
            t = time.time(); step_api(m, inputs, "nonexistent"); dt = time.time() - t
            line["cpu_baseline"] = {"value": f / dt, "unit": "steps/s", "cores": REF_THREADS, "host_cores": os.cpu_count(), "kind": "reference",
                                    "sample": f"first {nops} of {meta['ops']} graph ops ({100 * f:.1f}% of the UNet's Conv/MatMul FLOPs) through the reference's Model::run (fp16 weights, fp32 XNNPACK arithmetic), {dt:.1f} s, extrapolated to a full UNet run"}
        except Exception as e:   # the baseline must never take the bench down
            line["cpu_baseline"] = {"value": None, "unit": "steps/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"failed: {e}"}
    if RANK == 0:
        emit_line(line)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
