#!/usr/bin/env python
"""Diagnostic for the implicit-function backward: adjoint kernel vs finite differences of gclm_system, and the
per-component gradients of the solve vs finite differences of the solve."""
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import geocalib_b200 as G
from geocalib_b200.engine import get_engine, NUM_DIST
from test_gpu_parity import render_on_device, synth_params

KEYS = ("up_field", "latitude_field", "up_confidence", "latitude_confidence")
dev = torch.device("cuda:0")


def smooth(shape, gen):
    b, c, h, w = shape
    coarse = torch.rand((b, c, 5, 7), generator=gen) * 2 - 1
    return torch.nn.functional.interpolate(coarse, size=(h, w), mode="bilinear", align_corners=True).to(dev)


for model in sys.argv[1:] or ["pinhole", "simple_radial", "simple_divisional", "radial"]:
    B, h, w = 2, 96, 128
    K = NUM_DIST[model]; n = 3 + K
    cam_gt, grav_gt = synth_params(model, B, h, w, 91)
    f = render_on_device(model, cam_gt, grav_gt, h, w, dev, sigma=0.02, seed=13)
    gen = torch.Generator().manual_seed(5)
    D = {"up_field": smooth((B, 2, h, w), gen), "latitude_field": smooth((B, 1, h, w), gen),
         "up_confidence": 0.3 * smooth((B, 1, h, w), gen)[:, 0], "latitude_confidence": 0.3 * smooth((B, 1, h, w), gen)[:, 0]}
    opt = G.LMOptimizer({"camera_model": model, "num_steps": 40, "early_stop": False}).eval()
    eng = get_engine(dev)
    with torch.no_grad():
        out = opt(f)
    cam, grav = out["camera"]._data, out["gravity"]._data
    options = opt._options()
    Gs, Hs, _ = eng.system(model, *[f[k] for k in KEYS], cam, grav, options=options, rpf_basis=False)
    print(f"== {model}: |G| at the solution {Gs.abs().max().item():.3e}  diag H {Hs.diagonal(dim1=1, dim2=2)[0].tolist()}")
    # (1) adjoint kernel vs FD of the system in x
    lam = torch.randn((B, n), generator=gen).to(dev)
    ad = eng.adjoint(model, *[f[k] for k in KEYS], cam, grav, lam, options=options)
    for i, k in enumerate(KEYS):
        eps = 1e-2
        def lg(sign):
            d = {kk: (f[kk] + sign * eps * D[kk]).contiguous() if kk == k else f[kk] for kk in KEYS}
            Gx, _, _ = eng.system(model, *[d[kk] for kk in KEYS], cam, grav, options=options, rpf_basis=False)
            return (Gx.double() * lam.double()).sum(-1)
        fd = (lg(+1) - lg(-1)) / (2 * eps)
        an = (ad[i].double() * D[k].double()).flatten(1).sum(-1)
        print(f"  adjoint d(lam.G)/d{k:20s} FD {fd.tolist()}  kernel {an.tolist()}")
    # (2) per-component solve gradients
    comps = {"fy": lambda c, g: c[:, 3].sum(), "g0": lambda c, g: g[:, 0].sum(), "g1": lambda c, g: g[:, 1].sum(), "g2": lambda c, g: g[:, 2].sum()}
    if K: comps["k1"] = lambda c, g: c[:, 6].sum()
    if K == 2: comps["k2"] = lambda c, g: c[:, 7].sum()
    for cname, fn in comps.items():
        x = {k: f[k].clone().requires_grad_(True) for k in KEYS}
        c, g, _, _ = G.differentiable_calibrate(opt, x)
        fn(c, g).backward()
        line = f"  dL/dx, L={cname}: "
        for k in KEYS:
            eps = 2e-3 if "field" in k else 2e-2
            def sl(sign):
                d = {kk: (f[kk] + sign * eps * D[kk]).contiguous() if kk == k else f[kk] for kk in KEYS}
                with torch.no_grad():
                    o = opt(d)
                return float(fn(o["camera"]._data.double(), o["gravity"]._data.double()))
            fd = (sl(+1) - sl(-1)) / (2 * eps)
            an = float((x[k].grad.double() * D[k].double()).sum())
            line += f"{k[:6]}: FD {fd:+.4e} AN {an:+.4e} ({(an / fd - 1) * 100 if fd else 0:+.1f}%) | "
        print(line)
