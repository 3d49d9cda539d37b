#!/usr/bin/env python
"""Benchmark of the batched LM calibration path (BASELINE.json metric: images/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full LM calibration (num_steps LM iterations + the final cost/covariance pass) of
one batch of synthetic predicted fields per GPU.  Workload at every N (weak scaling, images are
independent problems -> no data-path collective): BASELINE.json configs[1],
batch=1024 pinhole 640x480, 20 LM iterations, early_stop off, per GPU.

Printed JSON (one line, rank 0):
  value      images/sec over all ranks, fields resident in HBM when the timed region starts
  e2e        same metric through the public API with HOST (pinned) buffers: H2D of the fields and
             D2H of the results inside the timed region, chunks pipelined on two streams
  roofline   fused pass kernel vs the measured HBM copy bandwidth (MEASURED_PEAKS.json);
             algorithmic bytes = B*H*W*5 planes*4 B per pass, passes = num_steps+1
  cpu_baseline  the CPU port of the reference algorithm (oracle/, OpenMP) on a bounded sample
`--impl reference` times that CPU port alone (the reference is PyTorch-eager Python and does not
travel to the GPU box; the oracle is pinned against it by tests/test_oracle_golden.py).
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (camera_model, batch per GPU, H, W, LM iterations)
    "pinhole_b1024": ("pinhole", 1024, 480, 640, 20),       # BASELINE.json configs[1] (default)
    "simple_radial_b4096": ("simple_radial", 4096, 480, 640, 20),  # configs[2]
    "divisional_b1024": ("simple_divisional", 1024, 1024, 1024, 20),  # configs[3] per-GPU share
    "radial_b1024": ("radial", 1024, 480, 640, 20),
    "pinhole_b1_320": ("pinhole", 1, 240, 320, 10),         # configs[0] shape (latency)
    "shared_b2048": ("pinhole", 2048, 480, 640, 20),        # configs[4] per-GPU share
}
PLANES = 5
BYTES_PER_PIXEL_PASS = PLANES * 4


def gt_params(model, B, h, w, seed):
    """Ground-truth cameras / gravities of the synthetic workload."""
    rng = np.random.default_rng(seed)
    roll = rng.uniform(-np.pi / 4, np.pi / 4, B)
    pitch = rng.uniform(-np.pi / 4, np.pi / 4, B)
    fish = model == "simple_divisional"
    vfov = np.deg2rad(rng.uniform(60, 120, B) if fish else rng.uniform(30, 90, B))
    fy = h / 2 / np.tan(vfov / 2)
    k1 = np.zeros(B)
    k2 = np.zeros(B)
    if model in ("simple_radial", "radial"):
        k1 = rng.uniform(-0.15, 0.15, B)
    if model == "radial":
        k2 = rng.uniform(-0.03, 0.03, B)
    if fish:
        k1 = rng.uniform(-0.4, -0.05, B)
    cam = np.stack([np.full(B, w), np.full(B, h), fy, fy, np.full(B, w / 2), np.full(B, h / 2), k1, k2], -1)
    sr, cr, sp, cp = np.sin(roll), np.cos(roll), np.sin(pitch), np.cos(pitch)
    grav = np.stack([-sr * cp, -cr * cp, sp], -1)
    return cam.astype(np.float32), grav.astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            if not (t0 - 0.05 <= t <= t1 + 0.15):
                continue
            p = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(p[1])); mx = float(p[2])
            except (ValueError, IndexError):
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[4:8]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak_gbs():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def ncu_traffic_per_launch(workload):
    """dram bytes per launch of the pass kernel from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "roofline_traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(workload)
        except Exception:
            return None
    return None


# --------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on the host cores (bounded sample of the same workload)
# --------------------------------------------------------------------------------------------------
def cpu_sample_fields(model, n_img, h, w, seed):
    from oracle import oracle

    cam, grav = gt_params(model, n_img, h, w, seed)
    rng = np.random.default_rng(seed + 1)
    ups, lats = [], []
    for b in range(n_img):
        u, l = oracle.render(model, cam[b], grav[b], h, w)
        ups.append(u); lats.append(l)
    up = np.stack(ups) + rng.normal(0, 0.02, (n_img, 2, h, w)).astype(np.float32)
    up = (up / np.linalg.norm(up, axis=1, keepdims=True)).astype(np.float32)
    lat = (np.stack(lats) + rng.normal(0, 0.02, (n_img, 1, h, w))).astype(np.float32)
    uc = rng.uniform(0.2, 1, (n_img, h, w)).astype(np.float32)
    lc = rng.uniform(0.2, 1, (n_img, h, w)).astype(np.float32)
    return up, lat, uc, lc


def cpu_run(model, fields, iters):
    from oracle import oracle

    t = time.perf_counter()
    oracle.lm(model, *fields, conf=dict(num_steps=iters, early_stop=False), compute_uncertainty=True)
    return time.perf_counter() - t


def host_threads():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_baseline(model, h, w, iters, target_seconds=15.0):
    cores = host_threads()
    probe_n = max(8, min(64, cores // 2))
    fields = cpu_sample_fields(model, probe_n, h, w, 99)
    dt = cpu_run(model, fields, iters)
    rate = probe_n / dt
    n = int(max(probe_n, min(1024, rate * target_seconds)))
    if n > probe_n:
        fields = cpu_sample_fields(model, n, h, w, 99)
        dt = cpu_run(model, fields, iters)
    else:
        n = probe_n
    return {"value": n / dt, "unit": "images/sec", "cores": cores, "kind": "port",
            "sample": f"{n} images of the workload ({h}x{w} {model}, {iters} LM iterations, early_stop off) "
                      f"in {dt:.1f} s; oracle/gclm_oracle.c, OpenMP over (image, row-chunk) tasks"}


def run_reference_arm(args, wl):
    """--impl reference: the CPU port of the reference path on all host cores; rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    model, B, h, w, iters = wl
    cores = host_threads()
    n = max(8, min(64, cores // 2))  # enough (image, row-chunk) tasks to occupy every host thread
    fields = cpu_sample_fields(model, n, h, w, 99)
    probe = cpu_run(model, fields, iters)
    # one step = a bounded sample: ~12 s of CPU work, at least 2 images per host thread so that the
    # (image, row-chunk) task pool keeps all threads busy (throughput still rises with sample size)
    per_step = int(min(512, max(n, 2 * cores if cores <= 128 else cores, (n / probe) * 12.0)))
    if per_step != n:
        fields = cpu_sample_fields(model, per_step, h, w, 99)
    for _ in range(min(args.warmup, 1)):
        cpu_run(model, fields, iters)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_run(model, fields, iters)
    dt = time.perf_counter() - t0
    val = per_step * args.steps / dt
    sample = (f"each step = {per_step} images of the workload on {cores} host threads "
              f"(oracle/gclm_oracle.c, CPU port of geocalib/lm_optimizer.py pinned by tests/golden)")
    line = {
        "impl": "reference", "metric": "images/sec (batched LM calibrate)", "value": val, "unit": "images/sec",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "camera_model": model, "batch_per_gpu": B, "height": h, "width": w,
                   "lm_iterations": iters, "early_stop": False},
        "cpu_baseline": {"value": val, "unit": "images/sec", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# --------------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pinhole_b1024", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="override images per GPU (debug)")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-chunk", type=int, default=128)
    ap.add_argument("--rows-per-tile", type=int, default=0, help="engine tiling knob (0 = auto)")
    ap.add_argument("--no-check", action="store_true", help="skip the calibration sanity check (experiments)")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.batch > 0:
        wl[1] = args.batch
    model, B, h, w, iters = wl
    if args.impl == "reference":
        run_reference_arm(args, wl)
        return

    import torch
    import torch.distributed as dist

    import geocalib_b200 as G
    from geocalib_b200.engine import get_engine

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    eng = get_engine(dev)
    shared = args.workload.startswith("shared")

    # ---- synthetic predicted fields, generated on this rank's device ----
    cam_gt, grav_gt = gt_params(model, B, h, w, 1234 + 1000 * 1 + rank)
    if shared:  # one camera for every image of every rank
        cam_gt[:, 2:4] = gt_params(model, 1, h, w, 4321)[0][0, 3]
    fields = eng.render_fields(model, torch.from_numpy(cam_gt).to(dev), torch.from_numpy(grav_gt).to(dev), h, w,
                               sigma=0.02, conf_lo=0.2, seed=1234 + rank)
    opt = G.LMOptimizer({"camera_model": model, "num_steps": iters, "early_stop": False,
                         "rows_per_tile": args.rows_per_tile}).eval()
    opt.shared_intrinsics = shared
    if shared and world > 1:
        from geocalib_b200.parallel import make_comm
        opt.comm = make_comm(dev)
    eng.reserve(B, h, w)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ----
    with torch.no_grad():
        for _ in range(args.warmup):
            out = opt(fields)
        barrier()
        sampler = ClockSampler(local)
        if rank == 0:
            sampler.start()
            time.sleep(0.25)
        l0 = eng.launch_count
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        eng.pass_timing(True)  # event pair around the pass launches of every timed solve (roofline)
        t_wall0 = time.time()
        ev0.record()
        for _ in range(args.steps):
            out = opt(fields)
        ev1.record()
        barrier()
        t_wall1 = time.time()
        pass_ms, pass_launches = eng.pass_time()
        eng.pass_timing(False)
        launches = eng.launch_count - l0
        ms = max_over_ranks(ev0.elapsed_time(ev1))
        clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    ms_per_step = ms / args.steps
    value = world * B * args.steps / (ms * 1e-3)

    # parity guard on the timed output: recovered focal close to the synthetic ground truth
    f_err = float((out["camera"]._data[:, 3].cpu() / torch.from_numpy(cam_gt[:, 3]) - 1).abs().median())
    assert args.no_check or f_err < 0.05, f"median focal error {f_err}: the timed run did not calibrate"

    # ---- roofline of the fused pass kernel ----
    passes = iters + 1
    bytes_per_launch = B * h * w * BYTES_PER_PIXEL_PASS
    # average duration of one pass launch inside the timed region: CUDA events recorded by the engine on
    # the solve stream around the `passes` back-to-back launches of each timed solve (max over ranks);
    # the once-per-solve set-up kernels (initial estimate, sin plane) are outside the bracket
    launch_ms = max_over_ranks(pass_ms / pass_launches) if pass_launches else ms_per_step / passes
    peak, peak_src = measured_peak_gbs()
    achieved = bytes_per_launch / (launch_ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_per_launch(args.workload), "peak_source": peak_src,
                "kernel": "gclm::lm_pass_kernel", "bytes_per_launch": bytes_per_launch,
                "launch_ms": launch_ms, "passes_per_step": passes, "step_ms_per_pass": ms_per_step / passes,
                "timed_launches": pass_launches}

    # ---- end to end: host (pinned) buffers -> H2D -> solve -> D2H, pipelined in chunks ----
    e2e = None
    if not args.no_e2e:
        from geocalib_b200.host_pipeline import HostPipeline
        chunk = min(args.e2e_chunk, B)
        pipe = HostPipeline(opt, B, h, w, chunk, dev)
        host_fields = {k: v.cpu().pin_memory() for k, v in fields.items()}
        with torch.no_grad():
            for _ in range(max(1, min(args.warmup, 2))):
                res = pipe.run(host_fields)
            barrier()
            t0 = time.perf_counter()
            for _ in range(args.steps):
                res = pipe.run(host_fields)
            barrier()
            dt = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([dt], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            dt = float(t.item())
        e2e = {"value": world * B * args.steps / dt, "unit": "images/sec",
               "h2d_bytes_per_step": pipe.h2d_bytes, "d2h_bytes_per_step": pipe.d2h_bytes,
               "chunk_images": chunk, "ms_per_step": 1e3 * dt / args.steps}
        assert torch.allclose(res["camera"][:, 3], out["camera"]._data[:, 3].cpu(), rtol=1e-4)
        del host_fields, pipe

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_baseline(model, h, w, iters)

    if rank == 0:
        line = {
            "metric": "images/sec (batched LM calibrate)", "value": value, "unit": "images/sec",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": args.workload, "camera_model": model, "batch_per_gpu": B, "global_batch": world * B,
                       "height": h, "width": w, "lm_iterations": iters, "early_stop": False,
                       "shared_intrinsics": shared, "parallelism": f"image-sharded x{world}, no data-path collective"
                       if not shared else f"image-sharded x{world}, one all-reduce per LM step",
                       "l2": f"inputs {B * h * w * BYTES_PER_PIXEL_PASS / 1e9:.2f} GB per GPU >> 126 MB L2 (no flush needed)"},
            "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches), "roofline": roofline,
            "cpu_baseline": cpu, "median_focal_error_vs_gt": f_err,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
