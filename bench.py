#!/usr/bin/env python3
"""bench.py — destripe throughput (Mpixel/s at 2048^2 uint16) on N B200s of one node, plus roofline and CPU baseline.

    python bench.py --gpus 1 --steps 8 --warmup 3                      # BASELINE.json configs[1] (the headline)
    python bench.py --config 3                                         # configs[2]: + Gaussian, 2x2 max, 8-bit output
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P bench.py --gpus N ...
    python bench.py --impl reference --steps 2 --warmup 1              # the reference's CPU path (oracle port) on the host cores

`--config K` selects BASELINE.json configs[K-1] (default 2 = the configuration the metric is quoted on):
    1  filter_streaks, sigma=(256,256), db10, level auto, wrap padding
    2  process_img / batch_filter: reflect padding, dark=100, flat-field
    3  config 2 + 5x5 Gaussian + 2x2 max down-sample + 16->8-bit right shift (uint8 output, 1/8 of the D2H bytes)
    4  lightsheet_correct background clean + destripe
    5  coif15, sigma=(128,512) (two passes, 90-tap filters)
One "step" = one pass over `--planes` planes per GPU.  Z planes are independent: with N GPUs every rank runs its own
planes (weak scaling, no collective on the data path).

One JSON line on stdout (rank 0).  Every GPU leg goes through the PUBLIC pystripe API (`core.process_img` /
`core.filter_streaks` / `core.batch_filter`):
  value             CUDA torch tensors in, CUDA tensors out (device-resident; CUDA events, max over ranks)
  e2e               numpy arrays over page-locked host memory (`core.pinned_empty`) in, host arrays out: H2D + kernels + D2H
  e2e_public_api    ordinary (pageable) numpy stack in, host arrays out
  e2e_batch_filter  `batch_filter` over uncompressed TIFF files on tmpfs -> TIFF files on tmpfs (decode, H2D, kernels, D2H, encode)
`roofline`: the dominant kernel (largest share of the step), algorithmic bytes / CUDA-event duration measured live,
against MEASURED_PEAKS.json.  `cpu_baseline`: the oracle port on the host cores (bounded sample, inputs pre-generated).
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path[:0] = [str(ROOT), str(ROOT / "image-preprocessing-pipeline_b200")]

import numpy as np  # noqa: E402

H = W = 2048
METRIC = "destripe_mpixel_per_s_2048x2048_uint16"
UNIT = "Mpixel/s"
_BASE = dict(sigma=(256, 256), wavelet="db10", level=0)
CONFIGS = {
    1: dict(fn="filter_streaks", flat=False, kw=dict(_BASE, padding_mode="wrap"),
            workload="configs[0]: filter_streaks, 2048x2048 uint16 planes, sigma=(256,256), db10, level auto, wrap pad"),
    2: dict(fn="process_img", flat=True, kw=dict(_BASE, padding_mode="reflect", dark=100),
            workload="configs[1]: batch_filter/process_img, 2048x2048 uint16 stack, sigma=(256,256), db10, level auto, reflect pad, dark=100, flat-field"),
    3: dict(fn="process_img", flat=True,
            kw=dict(_BASE, padding_mode="reflect", dark=100, gaussian_filter_2d=True, down_sample=(2, 2),
                    convert_to_8bit=True, bit_shift_to_right=8),
            workload="configs[2]: configs[1] + 5x5 Gaussian + 2x2 max down-sample + 16->8-bit right shift (uint8 out)"),
    4: dict(fn="process_img", flat=False, kw=dict(_BASE, padding_mode="wrap", lightsheet=True),
            workload="configs[3]: lightsheet_correct background clean + destripe (sigma=(256,256), db10, wrap pad), 2048x2048 uint16"),
    5: dict(fn="process_img", flat=False, kw=dict(sigma=(128, 512), wavelet="coif15", level=0, padding_mode="reflect"),
            workload="configs[4]: coif15, sigma=(128,512) (two passes), reflect pad, 2048x2048 uint16"),
}
N_DISTINCT = 8


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[K-1]")
    ap.add_argument("--planes", type=int, default=0, help="planes per step per GPU (default 250; 64 for configs 4, 5)")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("B200STRIPE_MAX_BATCH", "32")))
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-batch-filter", action="store_true", help="skip the file-based e2e leg")
    ap.add_argument("--cpu-planes", type=int, default=0, help="planes in the CPU sample (default: 2 per core; 1 per core for configs 4, 5)")
    ap.add_argument("--files", type=int, default=0, help="files in the batch_filter leg (default 1024 / ranks)")
    ap.add_argument("--fast", action="store_true", help="allow FMA contraction (exact=0); not the parity configuration")
    a = ap.parse_args()
    if a.planes <= 0:
        a.planes = 250 if a.config <= 3 else 64
    return a


def config_dict(a):
    """identical for both arms: the driver compares it"""
    return {"workload": CONFIGS[a.config]["workload"], "config_id": a.config, "plane": f"{H}x{W} uint16"}


# ------------------------------------------------------------------------------------------------ CPU arm (oracle)
def _cpu_init(config_id):
    """per worker process, untimed: the oracle, the flat-field and the N_DISTINCT synthetic input planes"""
    from oracle import pystripe_oracle as orc
    from tools import synth
    c = CONFIGS[config_id]
    _cpu_worker.orc = orc
    _cpu_worker.cfg = c
    _cpu_worker.flat = orc.normalize_flat(synth.flat_field((H, W))) if c["flat"] else None
    _cpu_worker.planes = [synth.plane(z, (H, W)) for z in range(N_DISTINCT)]


def _cpu_worker(z):
    orc, c = _cpu_worker.orc, _cpu_worker.cfg
    img = _cpu_worker.planes[z % N_DISTINCT].copy()
    t = time.perf_counter()
    if c["fn"] == "filter_streaks":
        out = orc.filter_streaks(img, **c["kw"])
    else:
        out = orc.process_img(img, flat=_cpu_worker.flat, **c["kw"])
    return time.perf_counter() - t, int(out[::64, ::64].astype(np.int64).sum())


def cpu_pool(config_id, cores):
    import multiprocessing as mp
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(config_id,))
    pool.map(_cpu_worker, range(cores))          # untimed: every worker has generated its inputs and compiled numba
    return pool


def cpu_throughput(config_id, n_planes, cores):
    """oracle port (reference algorithm on restated pywt) on `cores` processes, inputs pre-generated in every worker;
    returns (Mpx/s, seconds)."""
    with cpu_pool(config_id, cores) as pool:
        t = time.perf_counter()
        pool.map(_cpu_worker, range(n_planes), chunksize=1)
        dt = time.perf_counter() - t
    return n_planes * H * W / dt / 1e6, dt


def run_reference(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    cores = os.cpu_count() or 1
    per_step = a.cpu_planes or cores
    with cpu_pool(a.config, cores) as pool:
        for _ in range(max(a.warmup - 1, 0)):
            pool.map(_cpu_worker, range(cores), chunksize=1)
        t = time.perf_counter()
        for _ in range(a.steps):
            pool.map(_cpu_worker, range(per_step), chunksize=1)
        dt = time.perf_counter() - t
    v = a.steps * per_step * H * W / dt / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": dt / a.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": config_dict(a),
        "run": {"planes_per_step": per_step,
                "note": "CPU: reference algorithm (oracle port: PyWavelets restated in C, scipy.fftpack, glibc libm), one process "
                        "per host core like the reference's batch_filter farm; inputs pre-generated outside the timed region"},
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} planes of 2048x2048 per step, {a.steps} steps, inputs pre-generated"},
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
    from pystripe import core, _io
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

    def max_over_ranks(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cfg = CONFIGS[a.config]
    core.MAX_BATCH = a.batch
    core.EXACT = not a.fast
    P = a.planes
    flat = core.normalize_flat(synth.flat_field((H, W))) if cfg["flat"] else None

    def run(x):                                          # the public call, whatever x is (CUDA tensor / numpy)
        if cfg["fn"] == "filter_streaks":
            return core.filter_streaks(x, **cfg["kw"])
        return core.process_img(x, flat=flat, _max_batch=a.batch, **cfg["kw"])

    base = synth.stack(min(N_DISTINCT, P), (H, W), seed=1234 + 100 * rank)
    h_in = core.pinned_empty((P, H, W), np.uint16, device=local)     # page-locked host input of the e2e leg
    h_in[:] = np.concatenate([base] * (-(-P // base.shape[0])))[:P]
    d_in = torch.from_numpy(h_in).to(dev)

    # ---- device-resident: W warm-up, K timed steps, CUDA events on the launching (current) stream
    for _ in range(a.warmup):
        d_out = run(d_in)
    plan = next(reversed(core._plans.values()))          # the plan the public call built (geometry for the roofline model)
    ctx, info = plan.ctx, plan.info
    barrier()
    clocks = ClockSampler(local) if rank == 0 else None
    launches0 = ctx.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        d_out = run(d_in)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = ctx.launch_count - launches0
    clk = clocks.stop() if clocks else None
    ms_max = max_over_ranks(ms)
    value = world * a.steps * P * H * W / (ms_max * 1e-3) / 1e6
    out_plane_bytes = int(np.prod(d_out.shape[1:])) * d_out.element_size()

    def timed_host(fn, steps, warm):
        r = None
        for _ in range(max(2, warm)):      # two results are alive at a time (the old one while the new one is produced):
            r = fn()                       # both page-locked result blocks exist before the timed region starts
        barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            r = fn()
        torch.cuda.synchronize()
        return max_over_ranks(time.perf_counter() - t0), r

    # ---- end to end, public API, page-locked host input (H2D + kernels + D2H inside the timed region)
    dt, h_out = timed_host(lambda: run(h_in), a.steps, max(1, min(a.warmup, 2)))
    e2e_v = world * a.steps * P * H * W / dt / 1e6
    checksum = int(np.asarray(h_out)[:: max(1, P // 4), ::128, ::128].astype(np.int64).sum())
    assert np.array_equal(np.asarray(h_out)[:2], d_out[:2].cpu().numpy()), "host and device legs disagree"
    del h_out

    # ---- end to end, public API, ordinary pageable numpy stack
    pageable = np.array(h_in, copy=True)
    e2e_steps = max(2, a.steps // 2)
    dt, _ = timed_host(lambda: run(pageable), e2e_steps, 1)
    e2e_pageable = world * e2e_steps * P * H * W / dt / 1e6
    del pageable

    # ---- end to end through batch_filter over files on tmpfs (uncompressed TIFF in, TIFF out)
    bf = None
    if not a.no_batch_filter and cfg["fn"] == "process_img":
        try:
            bf = batch_filter_leg(a, cfg, core, _io, base, flat, local, world, barrier, max_over_ranks)
        except Exception as e:                                   # the file leg must never cost the bench line
            bf = {"error": f"{type(e).__name__}: {e}"}

    # ---- host<->device copy probe: the step's bytes with NO kernels, both directions at once, all ranks at the same time.
    # A probe of the box's PCIe / host-memory rate, NOT a ceiling (fewer, larger copies than the pipeline issues).
    copy_probe = None
    try:
        th_in = torch.from_numpy(h_in)
        h_o = core.pinned_empty((P,) + tuple(d_out.shape[1:]), core._native.CODE_TO_NP[info.out_dtype], device=local)
        th_out = torch.from_numpy(h_o)
        s_up, s_dn = torch.cuda.Stream(), torch.cuda.Stream()

        def copies():
            with torch.cuda.stream(s_up):
                d_in.copy_(th_in, non_blocking=True)
            with torch.cuda.stream(s_dn):
                th_out.copy_(d_out, non_blocking=True)
        copies()
        barrier()
        t0 = time.perf_counter()
        for _ in range(10):
            copies()
        torch.cuda.synchronize()
        dtc = max_over_ranks(time.perf_counter() - t0)
        copy_probe = {"value": world * 10 * P * H * W / dtc / 1e6, "unit": UNIT, "repetitions": 10,
                      "h2d_GBps_per_gpu": 10 * P * H * W * 2 / dtc / 1e9, "d2h_GBps_per_gpu": 10 * P * out_plane_bytes / dtc / 1e9,
                      "note": "same bytes per step moved H2D + D2H with no kernels, two large copies, all ranks at once; a probe, not a bound"}
    except Exception as e:
        print(f"copy probe failed: {type(e).__name__}: {e}", file=sys.stderr)

    # ---- roofline: per-launch CUDA events around every kernel (separate pass so `value` is not perturbed)
    roof, shares = None, {}
    if rank == 0:
        roof, shares = roofline(a, ctx, info, run, d_in, P, ms_max)

    # ---- CPU baseline (rank 0, N=1 only): the oracle port on the host cores, bounded sample
    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        cores = os.cpu_count() or 1
        n = a.cpu_planes or (2 * cores if a.config <= 3 else cores)
        v, dt_cpu = cpu_throughput(a.config, n, cores)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{n} planes of 2048x2048 (same workload, inputs pre-generated), {dt_cpu:.1f} s wall, one process per core"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_max / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": config_dict(a),
            "run": {"planes_per_step_per_gpu": P, "planes_per_launch": min(a.batch, P), "exact_summation_order": not a.fast,
                    "l2": "inputs larger than L2 (each step streams %d MB of uint16 input per GPU)" % (P * H * W * 2 // 2 ** 20),
                    "device_resident_streams": int(os.environ.get("B2S_DEV_SLOTS", "3")),
                    "parallelism": f"z-shard x{world}, no collective",
                    "api": "every GPU leg calls pystripe.core." + cfg["fn"] + " (value: CUDA tensors; e2e: pinned numpy; "
                           "e2e_public_api: pageable numpy; e2e_batch_filter: files)"},
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": int(P * H * W * 2),
                    "d2h_bytes_per_step": int(P * out_plane_bytes),
                    "timer": "host wall clock around the synchronous public call, max over ranks",
                    "input": "numpy array over page-locked memory (core.pinned_empty); output: page-locked array from the result pool"},
            "e2e_public_api": {"value": e2e_pageable, "unit": UNIT, "steps": e2e_steps,
                               "input": "pageable numpy stack (staged by the library's copy threads)"},
            "e2e_batch_filter": bf, "copy_only_probe": copy_probe,
            "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu, "clocks": clk,
            "kernel_time_shares": shares, "checksum": checksum,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def batch_filter_leg(a, cfg, core, _io, base, flat, local, world, barrier, max_over_ranks):
    """core.batch_filter over `n` uncompressed TIFF tiles on tmpfs -> TIFF tiles on tmpfs.  Every rank owns its own
    folder and GPU (B200STRIPE_DEVICES), so N ranks run N independent batch_filter calls at once."""
    n = a.files or max(128, 1024 // world)     # tmpfs holds the inputs and the outputs of every rank
    tmp_root = "/dev/shm" if os.path.isdir("/dev/shm") else None
    work = Path(tempfile.mkdtemp(prefix=f"b2s_bench_r{local}_", dir=tmp_root))
    saved = {k: os.environ.pop(k, None) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK")}   # one process = one independent farm
    os.environ["B200STRIPE_DEVICES"] = str(local)
    try:
        src, dst = work / "in", work / "out"
        src.mkdir()
        stack = np.concatenate([base] * (-(-n // base.shape[0])))[:n]
        _io.write_tiff_batch([src / f"img_{z:05d}.tif" for z in range(n)], np.ascontiguousarray(stack), None)
        cores = os.cpu_count() or 1
        kw = dict(cfg["kw"])
        if flat is not None:
            kw["flat"] = flat * 1.0            # batch_filter normalises: a normalised flat stays what it is
        res = {}
        for name, compression in (("uncompressed", None), ("adobe_deflate_1", ("ADOBE_DEFLATE", 1))):
            files = n
            flist = [src / f"img_{z:05d}.tif" for z in range(files)]

            def once():
                with open(os.devnull, "w") as null:
                    old = sys.stdout
                    sys.stdout = null
                    try:
                        rc = core.batch_filter(src, dst, files_list=flist, workers=max(2, cores // world), threads_per_gpu=8,
                                               compression=compression, **kw)
                    finally:
                        sys.stdout = old
                assert rc == 0, f"batch_filter returned {rc}"
            shutil.rmtree(dst, ignore_errors=True)
            once()                                               # warm-up: plan, pinned pools, the device encoder's buffers
            shutil.rmtree(dst, ignore_errors=True)               # (deleting the previous outputs is not part of the job)
            barrier()
            t0 = time.perf_counter()
            once()
            dt = max_over_ranks(time.perf_counter() - t0)
            res[name] = {"value": world * files * H * W / dt / 1e6, "unit": UNIT, "files": files, "seconds": dt}
        one = core.imread_tif_raw_png(dst / "img_00000.tif")
        res["output"] = f"{one.shape[0]}x{one.shape[1]} {one.dtype}"
        res["note"] = ("tmpfs -> native codec (libb2sio) -> pinned batch -> GPU -> pinned batch -> native codec -> tmpfs, "
                       f"{max(2, cores // world)} host threads; adobe_deflate_1 is the reference's default output compression: the strips are "
                       "deflated on the GPU (b2s_deflate_strips) and only compressed bytes cross PCIe and reach the files")
        return res
    finally:
        os.environ.pop("B200STRIPE_DEVICES", None)
        for k, v in saved.items():
            if v is not None:
                os.environ[k] = v
        shutil.rmtree(work, ignore_errors=True)


def notch_lane_ops_per_row(n):
    """FP32 lane-ops of the O(radix^2) phases of scipy's real transform of length n (rfftp, generic odd radices 7 .. 133),
    forward + backward, exact mode (separate multiply and add): per generic pass ido * l1 * (ipph - 1)^2 cos/sin pairs.
    Returns None for lengths that run as another plan class (even > 1000, Bluestein)."""
    if n < 2 or (n > 1000 and n % 2 == 0):
        return None
    f, m = [], n
    while m % 4 == 0:
        f.append(4); m //= 4
    if m % 2 == 0:
        f.append(2); m //= 2
    d = 3
    while d * d <= m:
        while m % d == 0:
            f.append(d); m //= d
        d += 2
    if m > 1:
        f.append(m)
    if any(p >= 135 for p in f):
        return None
    ops = 0
    for p in f:
        if p > 5:
            ipph = (p + 1) // 2
            ops += (n // p) * (ipph - 1) ** 2 * 2 * 2          # idl1 x pairs x (cos, sin) x (mul, add)
    return 2 * ops if ops else None


def roofline(a, ctx, info, run, d_in, P, ms_max):
    ctx.timing_enable(True)
    ctx.timing_read(reset=True)
    for _ in range(2):
        run(d_in)
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
    rows = [info.padded_height] + [info.level_rows[i] for i in range(info.levels)]
    cols = [info.padded_width] + [info.level_cols[i] for i in range(info.levels)]
    F = 90 if a.config == 5 else 20
    wh, ww = info.work_height, info.work_width

    def model(kname, lvl):
        """(algorithmic bytes, FP32 lane-ops in exact mode) per plane of one launch of this class at this level"""
        if kname in ("dwt_fwd", "dwt_inv") and lvl >= 1:
            b = 4 * rows[lvl - 1] * cols[lvl - 1] + 16 * rows[lvl] * cols[lvl]
            macs = 2 * F * rows[lvl] * cols[lvl - 1] + 4 * F * rows[lvl] * cols[lvl]
            return b, 2 * macs
        if kname == "notch" and lvl >= 1:
            per_row = notch_lane_ops_per_row(cols[lvl])
            return 8 * rows[lvl] * cols[lvl], (per_row * rows[lvl] if per_row else None)
        if kname == "prologue":
            return 2 * wh * ww + (4 * wh * ww if CONFIGS[a.config]["flat"] and a.config == 2 else 0) + 4 * rows[0] * cols[0], None
        if kname == "epilogue":
            return 4 * wh * ww + int(info.out_height * info.out_width * (1 if info.out_dtype == 0 else 2)), None
        return int(info.algorithmic_bytes_per_plane), None

    per_plane, lane_ops = model(kname, lvl)
    planes_timed = 2 * P
    n_pass = max(1, info.n_passes)
    avg_launch_ms = kms / kn
    launches_per_plane = kn / planes_timed
    alg_bytes_per_launch = per_plane * n_pass / launches_per_plane if kname in ("dwt_fwd", "dwt_inv", "notch") else per_plane / launches_per_plane
    achieved = alg_bytes_per_launch / (avg_launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    for tf in (ROOT / "profiles" / "r02_traffic.json", ROOT / "profiles" / "r01_traffic.json"):
        if tf.exists() and a.config == 2:
            tj = json.loads(tf.read_text())
            ent = tj["kernels"].get(f"{kname}@L{lvl}" if lvl else kname)
            if ent:
                traffic = ent["dram_bytes_per_plane"] * n_pass / launches_per_plane
                traffic_src = tj["source"]
                break
    fp32_peak = 148 * 128 * 1.965e9
    itemsize = {0: 1, 1: 2, 2: 4}
    io_bytes = H * W * 2 + info.out_height * info.out_width * itemsize.get(int(info.out_dtype), 2)
    fp32 = {"flops_per_plane_model": int(info.flops_per_plane), "peak_lane_ops_per_s": fp32_peak,
            "whole_pipeline_frac": info.flops_per_plane * a.steps * P / (ms_max * 1e-3) / fp32_peak,
            "kernel_frac": (lane_ops * n_pass * planes_timed / (kms * 1e-3) / fp32_peak) if lane_ops else None,
            "note": "exact mode issues a separate multiply and add per tap / butterfly term (2 lane-ops per MAC): instruction issue "
                    "and the FP32 pipe, not HBM, bound the DWT and notch kernels; kernel_frac = the dominant kernel's modelled "
                    "lane-ops (DWT: taps; notch: the O(radix^2) phases of scipy's generic radices only) / its time / peak"}
    roof = {"bound": "hbm", "binding_in_practice": "instruction issue / FP32 pipe (ncu: profiles/r02_final_main_kernels.md)",
            "kernel": f"{kname}@level{lvl}" if lvl else kname, "achieved": achieved, "peak": peak,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "fp32": fp32,
            "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 (of fallback)",
            "algorithmic_bytes_per_launch": alg_bytes_per_launch, "avg_launch_ms": avg_launch_ms,
            "planes_per_launch": min(a.batch, P), "share_of_step": round(kms / total, 4), "largest_mover": None,
            "whole_pipeline": {"algorithmic_bytes_per_plane": int(info.algorithmic_bytes_per_plane),
                               "achieved_GBps": info.algorithmic_bytes_per_plane * a.steps * P / (ms_max * 1e-3) / 1e9,
                               "frac": info.algorithmic_bytes_per_plane * a.steps * P / (ms_max * 1e-3) / 1e9 / peak},
            # the same run against the bytes that MUST cross HBM (input planes once, output planes once)
            "compulsory_io": {"bytes_per_plane": int(io_bytes),
                              "achieved_GBps": io_bytes * a.steps * P / (ms_max * 1e-3) / 1e9,
                              "frac": io_bytes * a.steps * P / (ms_max * 1e-3) / 1e9 / peak,
                              "note": "input + output planes only; the stage model above counts every kernel's reads and writes"}}
    if ("dwt_fwd", 1) in tm:   # the kernel that moves the most bytes, for the HBM view of the same run
        kms1, kn1 = tm[("dwt_fwd", 1)]
        b1, ops1 = model("dwt_fwd", 1)
        ach1 = b1 * n_pass * planes_timed / (kms1 * 1e-3) / 1e9
        roof["largest_mover"] = {"kernel": "dwt_fwd@level1", "achieved": ach1, "frac": ach1 / peak,
                                 "fp32_frac": ops1 * n_pass * planes_timed / (kms1 * 1e-3) / fp32_peak,
                                 "algorithmic_bytes_per_plane": b1, "share_of_step": round(kms1 / total, 4)}
    return roof, shares


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
