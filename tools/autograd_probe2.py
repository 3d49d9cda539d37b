#!/usr/bin/env python
"""Diagnostic: dG/dtheta by finite differences vs -H, and convergence of the solve, for one model / Huber scale."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import geocalib_b200 as G
from geocalib_b200 import autograd as AG
from geocalib_b200.engine import get_engine, NUM_DIST
from test_gpu_parity import render_on_device, synth_params
from test_gpu_autograd import smooth_direction, KEYS

dev = torch.device("cuda:0")
torch.set_printoptions(precision=4, linewidth=200, sci_mode=False)
a = float(sys.argv[2]) if len(sys.argv) > 2 else 0.2
for model in sys.argv[1].split(","):
    B, h, w = 2, 96, 128
    K = NUM_DIST[model]; n = 3 + K
    cam_gt, grav_gt = synth_params(model, B, h, w, 91)
    print("==", model, "gt cam", cam_gt[:, [3, 6, 7]].tolist(), "a", a)
    f = render_on_device(model, cam_gt, grav_gt, h, w, dev, sigma=0.02, seed=13)
    gen = torch.Generator().manual_seed(5)
    D = {"up_field": smooth_direction((B, 2, h, w), gen, dev), "latitude_field": smooth_direction((B, 1, h, w), gen, dev)}
    eng = get_engine(dev)
    sols = {}
    for steps in (50, 200):
        opt = G.LMOptimizer({"camera_model": model, "num_steps": steps, "early_stop": False, "up_loss_fn_scale": a, "lat_loss_fn_scale": a}).eval()
        with torch.no_grad():
            out = opt(f)
        sols[steps] = (out["camera"]._data.clone(), out["gravity"]._data.clone())
        Gs, Hs, _ = eng.system(model, *[f[k] for k in KEYS], *sols[steps], options=opt._options(), rpf_basis=False)
        print(f" steps {steps}: f,k = {out['camera']._data[:, [3, 6, 7]].tolist()}  G = {Gs.tolist()}  stop_at {out['stop_at'].tolist()}")
    opt = G.LMOptimizer({"camera_model": model, "num_steps": 50, "early_stop": False, "up_loss_fn_scale": a, "lat_loss_fn_scale": a}).eval()
    cam, grav = sols[50]
    options = opt._options()
    Gs, Hs, _ = eng.system(model, *[f[k] for k in KEYS], cam, grav, options=options, rpf_basis=False)
    for hstep in (3e-4, 1e-3, 3e-3):
        A = torch.zeros((B, n, n), dtype=torch.float64, device=dev)
        for j in range(n):
            def G_at(s):
                d = torch.zeros((B, n), device=dev); d[:, j] = s * hstep
                gr = AG.sphere_plus(grav, d[:, :2]); cm = cam.clone(); cm[:, 2:4] = cam[:, 2:4] * torch.exp(d[:, 2:3])
                if K >= 1: cm[:, 6] = cam[:, 6] + d[:, 3]
                if K == 2: cm[:, 7] = cam[:, 7] + d[:, 4]
                elif K == 1: cm[:, 7] = cm[:, 6]
                return eng.system(model, *[f[k] for k in KEYS], cm.contiguous(), gr.contiguous(), options=options, rpf_basis=False)[0].double()
            A[:, :, j] = (G_at(+1) - G_at(-1)) / (2 * hstep)
        print(f" h={hstep}: -A[0] =\n{(-A[0]).cpu()}")
    print(f" H[0] =\n{Hs[0].cpu()}")
    # gradient per component at two FD steps
    comps = {"fy": lambda c, g: c[:, 3].sum() / 100, "k1": lambda c, g: c[:, 6].sum(), "g1": lambda c, g: g[:, 1].sum()}
    for cname, fn in comps.items():
        x = {k: f[k].clone().requires_grad_(True) for k in KEYS}
        c, g, _, _ = G.differentiable_calibrate(opt, x)
        fn(c, g).backward()
        line = f"  L={cname}: "
        for k in ("up_field", "latitude_field"):
            an = float((x[k].grad.double() * D[k].double()).sum())
            line += f"{k[:3]} AN {an:+.5e}"
            for eps in (1e-3, 4e-3):
                def sl(sign):
                    d = {kk: (f[kk] + sign * eps * D[kk]).contiguous() if kk == k else f[kk] for kk in KEYS}
                    with torch.no_grad():
                        o = opt(d)
                    return float(fn(o["camera"]._data.double(), o["gravity"]._data.double()))
                line += f" FD({eps:g}) {(sl(+1) - sl(-1)) / (2 * eps):+.5e}"
            line += " | "
        print(line)
