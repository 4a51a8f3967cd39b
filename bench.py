#!/usr/bin/env python3
"""bench.py — destripe throughput (Mpixel/s at 2048^2 uint16) on N B200s of one node, plus roofline and CPU baseline.

    python bench.py --gpus 1 --steps 8 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --steps 2 --warmup 1        # the reference's CPU path (oracle port) on the host cores

Workload (BASELINE.json configs[1]): process_img / batch_filter semantics over a synthetic 2048x2048 uint16 tile stack,
sigma=(256,256), wavelet db10, level auto, padding 'reflect', dark=100, flat-field; one "step" = one pass over
`--planes` planes (default 250, so the default 8 steps cover the 2000-plane stack).  Z planes are independent: with
N GPUs every rank runs its own `--planes` planes per step (weak scaling, no collective on the data path).

One JSON line on stdout (rank 0).  `value`: device-resident throughput (CUDA events, max over ranks).  `e2e`: the same
call with pinned HOST buffers through the C ABI (H2D + kernels + D2H inside the timed region).  `roofline`: the dominant
kernel (largest share of the step), algorithmic bytes / CUDA-event duration measured live, against MEASURED_PEAKS.json.
`cpu_baseline`: the oracle port on the host cores (bounded sample).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]

import numpy as np  # noqa: E402

H = W = 2048
WORK = dict(sigma=(256, 256), wavelet="db10", level=0, padding_mode="reflect", dark=100)
WORKLOAD = "configs[1]: batch_filter/process_img, 2048x2048 uint16 stack, sigma=(256,256), db10, level auto, reflect pad, dark=100, flat-field"
METRIC = "destripe_mpixel_per_s_2048x2048_uint16"
UNIT = "Mpixel/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--planes", type=int, default=250, help="planes per step per GPU")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("B200STRIPE_MAX_BATCH", "32")))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic planes (tiled to --planes)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-planes", type=int, default=0, help="planes in the CPU sample (default: 2 per core)")
    ap.add_argument("--fast", action="store_true", help="allow FMA contraction (exact=0); not the parity configuration")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def _cpu_worker(args):
    z, = args
    from oracle import pystripe_oracle as orc
    from tools import synth
    img = synth.plane(z % 8, (H, W))
    flat = _cpu_worker.flat
    t = time.perf_counter()
    out = orc.process_img(img, flat=flat, **WORK)
    return time.perf_counter() - t, int(out[::64, ::64].sum())


def _cpu_init():
    from oracle import pystripe_oracle as orc
    from tools import synth
    _cpu_worker.flat = orc.normalize_flat(synth.flat_field((H, W)))


def cpu_throughput(n_planes, cores):
    """oracle port (reference algorithm on restated pywt) on `cores` processes; returns (Mpx/s, seconds)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        pool.map(_cpu_worker, [(z,) for z in range(cores)])          # warm-up (numba / page-in), untimed
        t = time.perf_counter()
        pool.map(_cpu_worker, [(z,) for z in range(n_planes)])
        dt = time.perf_counter() - t
    return n_planes * H * W / dt / 1e6, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = a.cpu_planes or cores
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    with ctx.Pool(cores, initializer=_cpu_init) as pool:
        for _ in range(max(a.warmup, 1)):
            pool.map(_cpu_worker, [(z,) for z in range(cores)])
        t = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_cpu_worker, [(z,) for z in range(per_step)])
        dt = time.perf_counter() - t
    v = a.steps * per_step * H * W / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": WORKLOAD, "planes_per_step": per_step, "note": "CPU: reference algorithm (oracle port; PyWavelets restated in C, scipy.fftpack, glibc libm), one process per host core"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} planes of 2048x2048 per step, {a.steps} steps"},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.rows = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for nme, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------ GPU arm
def run_b200(a):
    import torch
    import torch.distributed as dist
    from pystripe import core, _native
    from tools import synth

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    P = a.planes
    base = synth.stack(min(a.distinct, P), (H, W), seed=1234 + 100 * rank)
    flat = core.normalize_flat(synth.flat_field((H, W)))
    plan = core._get_plan(local, (H, W), _native.U16, process=1, threshold=None, bidirectional=False, log1p=True,
                          flat=flat, out_code=_native.U16, max_batch=a.batch, exact=0 if a.fast else 1, **WORK)
    ctx = plan.ctx
    info = plan.info
    # pinned host buffers for the end-to-end leg, device-resident copies for the kernel leg
    h_in = ctx.pinned_empty((P, H, W), np.uint16)
    h_out = ctx.pinned_empty((P,) + plan.out_shape, plan.out_dtype)
    reps = -(-P // base.shape[0])
    h_in[:] = np.concatenate([base] * reps)[:P]
    d_in = torch.from_numpy(h_in).to(dev)
    d_out = torch.empty((P,) + plan.out_shape, dtype=torch.uint16, device=dev)

    # ---- device-resident: W warm-up, K timed steps, CUDA events on the launching (current) stream
    for _ in range(a.warmup):
        plan.run_torch(d_in, d_out)
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        plan.run_torch(d_in, d_out)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clk = clocks.stop() if clocks else None
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * a.steps * P * H * W / (ms_max * 1e-3) / 1e6

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H inside the timed region)
    for _ in range(max(1, min(a.warmup, 2))):
        plan.run_host(h_in, h_out)
    barrier()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        plan.run_host(h_in, h_out)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_v = world * a.steps * P * H * W / float(t.item()) / 1e6
    checksum = int(h_out[:: max(1, P // 4), ::128, ::128].astype(np.int64).sum())

    # ---- the platform's ceiling for e2e: one step's host<->device bytes with NO kernels, both directions at once, all
    # ranks at the same time (pinned buffers of the e2e leg, one stream per direction).  Not a bench value.
    copy_only = None
    try:
        th_in, th_out = torch.from_numpy(h_in), torch.from_numpy(h_out)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def copies():
            with torch.cuda.stream(s_up):
                d_in.copy_(th_in, non_blocking=True)
            with torch.cuda.stream(s_dn):
                th_out.copy_(d_out, non_blocking=True)
        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(3):
            copies()
        torch.cuda.synchronize()
        t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copy_only = world * 3 * P * H * W / float(t.item()) / 1e6
    except Exception as e:                                      # the probe must never cost the bench line
        print(f"copy-only probe failed: {type(e).__name__}: {e}", file=sys.stderr)

    # ---- roofline: per-launch CUDA events around every kernel (separate pass so `value` is not perturbed)
    roof = None
    shares = {}
    if rank == 0:
        ctx.timing_enable(True)
        ctx.timing_read(reset=True)
        for _ in range(2):
            plan.run_torch(d_in, d_out)
        tm = ctx.timing_read(reset=True, per_level=True)
        ctx.timing_enable(False)
        total = sum(v[0] for v in tm.values())
        shares = {f"{k}@L{l}" if l else k: round(v[0] / total, 4) for (k, l), v in sorted(tm.items(), key=lambda kv: -kv[1][0])}
        (kname, lvl), (kms, kn) = max(tm.items(), key=lambda kv: kv[1][0])
        peaks = {}
        pf = ROOT / "MEASURED_PEAKS.json"
        if pf.exists():
            peaks = json.loads(pf.read_text())
        peak = float(peaks.get("hbm_gbs", 6650.0))
        nb_per_launch = min(a.batch, P)
        rows = [info.padded_height] + [info.level_rows[i] for i in range(info.levels)]
        cols = [info.padded_width] + [info.level_cols[i] for i in range(info.levels)]
        if kname in ("dwt_fwd", "dwt_inv") and lvl >= 1:
            per_plane = 4 * rows[lvl - 1] * cols[lvl - 1] + 16 * rows[lvl] * cols[lvl]
        elif kname == "notch" and lvl >= 1:
            per_plane = 8 * rows[lvl] * cols[lvl]
        elif kname == "prologue":
            per_plane = 2 * H * W + 4 * H * W + 4 * rows[0] * cols[0]
        elif kname == "epilogue":
            per_plane = 4 * H * W + 2 * H * W
        else:
            per_plane = info.algorithmic_bytes_per_plane
        planes_timed = 2 * P
        avg_launch_ms = kms / kn
        launches_per_plane = kn / planes_timed
        alg_bytes_per_launch = per_plane / launches_per_plane
        achieved = alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
        # DRAM traffic of that kernel from the committed `ncu --set full` capture (dram__bytes_read+write per launch)
        traffic, traffic_src = None, None
        tf = ROOT / "profiles" / "r01_traffic.json"
        if tf.exists():
            tj = json.loads(tf.read_text())
            ent = tj["kernels"].get(f"{kname}@L{lvl}" if lvl else kname)
            if ent:
                traffic = ent["dram_bytes_per_plane"] / launches_per_plane
                traffic_src = tj["source"]
        # the competing ceiling: FP32 lane-ops (exact mode issues a separate multiply and add per tap)
        fp32_peak = 148 * 128 * 1.965e9
        fp32 = {"flops_per_plane_model": int(info.flops_per_plane), "peak_lane_ops_per_s": fp32_peak,
                "whole_pipeline_frac": info.flops_per_plane * a.steps * P / (ms_max * 1e-3) / fp32_peak,
                "note": "DWT multiply-adds only (2 lane-ops per MAC in exact mode); FFT, log1p/expm1 and index work come on top"}
        roof = {"bound": "hbm", "kernel": f"{kname}@level{lvl}" if lvl else kname, "achieved": achieved, "peak": peak,
                "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "fp32": fp32,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
                "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_launch_ms,
                "planes_per_launch": nb_per_launch, "share_of_step": round(kms / total, 4),
                "largest_mover": None,
                "whole_pipeline": {"algorithmic_bytes_per_plane": int(info.algorithmic_bytes_per_plane),
                                   "achieved_GBps": info.algorithmic_bytes_per_plane * a.steps * P / (ms_max * 1e-3) / 1e9,
                                   "frac": info.algorithmic_bytes_per_plane * a.steps * P / (ms_max * 1e-3) / 1e9 / peak}}

    # the kernel that moves the most bytes (forward DWT level 1), for the HBM view of the same run
    if roof is not None and ("dwt_fwd", 1) in tm:
        kms1, kn1 = tm[("dwt_fwd", 1)]
        per_plane1 = 4 * rows[0] * cols[0] + 16 * rows[1] * cols[1]
        ach1 = per_plane1 * (2 * P) / (kms1 * 1e-3) / 1e9
        roof["largest_mover"] = {"kernel": "dwt_fwd@level1", "achieved": ach1, "frac": ach1 / peak,
                                 "algorithmic_bytes_per_plane": per_plane1, "share_of_step": round(kms1 / total, 4)}

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_planes or 2 * cores
        v, dt_cpu = cpu_throughput(n, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} planes of 2048x2048 (same workload), {dt_cpu:.1f} s wall, one process per core"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": WORKLOAD, "planes_per_step_per_gpu": P, "planes_per_launch": min(a.batch, P),
                       "exact_summation_order": not a.fast,
                       "l2": "inputs larger than L2 (each step streams %d MB of uint16 input per GPU)" % (P * H * W * 2 // 2 ** 20),
                       "device_resident_streams": int(os.environ.get("B2S_DEV_SLOTS", "3")),
                       "parallelism": f"z-shard x{world}, no collective"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(P * H * W * 2),
                    "d2h_bytes_per_step": int(P * plan.out_shape[0] * plan.out_shape[1] * np.dtype(plan.out_dtype).itemsize),
                    "timer": "host wall clock around the synchronous C-ABI call (internal streams), max over ranks",
                    "copy_only_ceiling": copy_only,
                    "frac_of_copy_only_ceiling": (e2e_v / copy_only) if copy_only else None,
                    "copy_only_note": "same bytes per step moved H2D + D2H with no kernels, all ranks at once (PCIe / host "
                                      "memory ceiling of this box, in the metric's unit); e2e / ceiling is the overlap achieved"},
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
            "kernel_time_shares": shares, "checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def _claim_stdout():
    """keep fd 1 for the single JSON line: libraries (NCCL prints its version banner on stdout) write to stderr instead."""
    real = os.dup(1)
    os.dup2(2, 1)
    sys.stdout = os.fdopen(real, "w", buffering=1)


def main():
    a = parse()
    _claim_stdout()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
