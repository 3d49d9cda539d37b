"""Measures how far the engine's covariance / uncertainties are from the reference goldens (GPU).
Used to set (and justify) the tolerance in tests/test_gpu_parity.py::test_golden_lm."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
from test_gpu_parity import GOLDEN_CASES, _load, run_engine_on_golden  # noqa: E402

dev = torch.device("cuda", 0)
gd = ROOT / "tests" / "golden"
worst = {}
for case in GOLDEN_CASES:
    g = _load(gd, f"lm_{case}.npz")
    out = run_engine_on_golden(g, dev)
    cov, cr = out["covariance"].cpu().numpy().astype(np.float64), g["out_covariance"].astype(np.float64)
    sd = np.sqrt(np.abs(np.einsum("bii->bi", cr)))
    e_cov = (np.abs(cov - cr) / (sd[:, :, None] * sd[:, None, :])).max()
    row = {"cov_scaled": e_cov, "cond": float(np.linalg.cond(cr / (sd[:, :, None] * sd[:, None, :])).max())}
    for k in ("roll_uncertainty", "pitch_uncertainty", "gravity_uncertainty", "focal_uncertainty", "vfov_uncertainty"):
        a, b = out[k].cpu().numpy().astype(np.float64), g["out_" + k].astype(np.float64)
        row[k] = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-12))) if np.any(b != 0) else 0.0
    dp = np.abs(out["camera"]._data.cpu().numpy()[:, 3] / g["out_camera"][:, 3] - 1).max()
    row["focal_rel"] = float(dp)
    print(case, {k: f"{v:.2e}" for k, v in row.items()})
    for k, v in row.items():
        worst[k] = max(worst.get(k, 0), v)
print("WORST", {k: f"{v:.2e}" for k, v in worst.items()})
