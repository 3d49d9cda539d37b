#!/usr/bin/env python
"""Where the reference CNN's GPU time goes (one 320x416 input, eager, fp32): kernels aggregated by name.
usage: frontend_profile.py [ncu]   -- with 'ncu' it only runs 1 warm-up + 2 forwards (to be wrapped by ncu)."""
import sys
import tempfile
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "baseline"))
import ref_import

ref_import.import_reference()
import geocalib
from geocalib.geocalib import GeoCalib as Net

from geocalib_b200 import frontend as FE

torch.manual_seed(7)
with tempfile.TemporaryDirectory() as d:
    torch.save({"model": Net().state_dict()}, Path(d) / "w.tar")
    ex = geocalib.GeoCalib(weights=str(Path(d) / "w.tar")).cuda().eval()
x = torch.rand(1, 3, 320, 416, generator=torch.Generator().manual_seed(3)).cuda()
FE.use_device_bases(ex.model)
with torch.no_grad():
    for _ in range(3 if len(sys.argv) < 2 else 1):
        FE.cnn_forward(ex.model, x, capturable=True)
    torch.cuda.synchronize()
    if len(sys.argv) > 1:
        for _ in range(2):
            FE.cnn_forward(ex.model, x, capturable=True)
        torch.cuda.synchronize()
        sys.exit(0)
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(5):
            FE.cnn_forward(ex.model, x, capturable=True)
        torch.cuda.synchronize()
    ev = prof.key_averages()
    rows = sorted(((e.device_time_total / 5, e.count / 5, e.key) for e in ev if e.device_time_total > 0), reverse=True)
    total = sum(r[0] for r in rows)
    print(f"total device time per forward: {total / 1e3:.3f} ms over {sum(r[1] for r in rows):.0f} kernels")
    for t, c, k in rows[:25]:
        print(f"{t / 1e3:8.3f} ms {100 * t / total:5.1f}% x{c:6.1f}  {k[:150]}")
