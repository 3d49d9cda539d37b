"""Tiny end-to-end solves for compute-sanitizer (memcheck / racecheck / initcheck / synccheck)."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import geocalib_b200 as G
from geocalib_b200.engine import get_engine

dev = torch.device("cuda", 0)
eng = get_engine(dev)
for model, shared in (("pinhole", False), ("radial", False), ("simple_divisional", False), ("pinhole", True)):
    B, h, w = 3, 40, 160
    fy = np.full(B, h / 2 / np.tan(np.deg2rad(60) / 2))
    k1 = np.full(B, -0.1 if model != "pinhole" else 0.0)
    cam = np.stack([np.full(B, w), np.full(B, h), fy, fy, np.full(B, w / 2), np.full(B, h / 2), k1, 0 * k1], -1).astype(np.float32)
    grav = np.tile(np.array([[0.1, -0.95, 0.2]], np.float32), (B, 1))
    grav /= np.linalg.norm(grav, axis=-1, keepdims=True)
    f = eng.render_fields(model, torch.from_numpy(cam).to(dev), torch.from_numpy(grav).to(dev), h, w, sigma=0.02, conf_lo=0.2, seed=3)
    # B = 3 fits the persistent schedule (one cooperative launch, grid barriers); per_pass = one launch per pass;
    # shared: the fused tail (reduce + update + stop latch in the last CTA)
    for sched in (("per_pass", "auto") if not shared else ("auto",)):
        opt = G.LMOptimizer({"camera_model": model, "num_steps": 3, "early_stop": False, "schedule": sched}).eval()
        opt.shared_intrinsics = shared
        out = opt(f)
        out2 = opt({"up_field": f["up_field"], "latitude_field": f["latitude_field"]})  # generic (no-confidence) path
        torch.cuda.synchronize()
        print(model, shared, sched, out["camera"]._data[:, 3].tolist(), out2["camera"]._data[:, 3].tolist())
    if not shared:  # pre-transformed layout with fp16 confidences (cp.async 8-byte copies)
        out3 = G.LMOptimizer({"camera_model": model, "num_steps": 3, "early_stop": False, "confidence_dtype": "float16"}).eval()(f)
        torch.cuda.synchronize()
        print(model, "conf16", out3["camera"]._data[:, 3].tolist())
print("done")
