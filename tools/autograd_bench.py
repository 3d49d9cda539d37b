#!/usr/bin/env python
"""Cost of the differentiable solve at the BASELINE image size: forward (fused solve), backward (2(3+K) system passes
+ the adjoint pass), and the adjoint kernel alone against the HBM roofline (20 B read + 20 B written per pixel)."""
import json
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import geocalib_b200 as G
from geocalib_b200.engine import NUM_DIST, get_engine
from oracle import oracle

KEYS = ("up_field", "latitude_field", "up_confidence", "latitude_confidence")
dev = torch.device("cuda:0")
peaks = json.load(open(ROOT / "MEASURED_PEAKS.json")) if (ROOT / "MEASURED_PEAKS.json").exists() else {}


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n


for model, B in (("pinhole", 512), ("simple_radial", 512)):
    h, w = 480, 640
    cam_gt, grav_gt = oracle.gt_params(model, B, h, w, 77)
    eng = get_engine(dev)
    f = eng.render_fields(model, torch.from_numpy(cam_gt).to(dev), torch.from_numpy(grav_gt).to(dev), h, w, sigma=0.02, conf_lo=0.2, seed=1)
    opt = G.LMOptimizer({"camera_model": model, "num_steps": 30}).eval()
    with torch.no_grad():
        out = opt(f)
    cam, grav = out["camera"]._data, out["gravity"]._data
    lam = torch.randn((B, 3 + NUM_DIST[model]), device=dev)
    options = opt._options()
    t_fwd = timed(lambda: opt(f), 3)
    t_adj = timed(lambda: eng.adjoint(model, *[f[k] for k in KEYS], cam, grav, lam, options=options), 10)
    x = {k: f[k].clone().requires_grad_(True) for k in KEYS}

    def fwd_bwd():
        for k in KEYS:
            x[k].grad = None
        c, g, _, _ = G.differentiable_calibrate(opt, x)
        (c[:, 3].sum() / 100 + g[:, 1].sum()).backward()

    t_fb = timed(fwd_bwd, 3)
    gb = B * h * w * 40 / 1e9
    rec = {"model": model, "batch": B, "size": [h, w], "iterations_run": int(out["stop_at"].max()), "forward_ms": round(t_fwd, 3),
           "forward_backward_ms": round(t_fb, 3), "backward_ms": round(t_fb - t_fwd, 3), "adjoint_kernel_ms": round(t_adj, 3),
           "adjoint_GBps": round(gb / (t_adj * 1e-3), 1), "hbm_peak_GBps": peaks.get("hbm_gbs")}
    print(json.dumps(rec))
