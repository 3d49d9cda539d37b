"""Where does a pass go for ONE small image?  Needs the -DGCLM_PHASE_TIMERS build (GCLM_LIB=tools/variants/timers.so).
Prints, per pass, clock64 deltas (in us at the measured SM clock) of the phases of the CTA that ran the solver."""
import ctypes, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import geocalib_b200 as G
from geocalib_b200.engine import get_engine
from oracle import oracle

dev = torch.device("cuda", 0)
model, B, h, w, iters = "pinhole", 1, 240, 320, 10
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 0
sched = sys.argv[2] if len(sys.argv) > 2 else "persistent"
cam, grav = oracle.gt_params(model, B, h, w, 5)
eng = get_engine(dev)
f = eng.render_fields(model, torch.from_numpy(cam).to(dev), torch.from_numpy(grav).to(dev), h, w, sigma=0.02, conf_lo=0.2, seed=3)
opt = G.LMOptimizer({"camera_model": model, "num_steps": iters, "early_stop": False, "rows_per_tile": rows, "schedule": sched}).eval()
with torch.no_grad():
    for _ in range(3):
        opt(f)
torch.cuda.synchronize()
cfg = (ctypes.c_int32 * 5)()
eng.lib.gclm_tile_config(B, h, w, 0, 148, rows, cfg)
ncta = B * cfg[3]
n = 400000 + (iters + 1) * 8
buf = (ctypes.c_longlong * n)()
eng.lib.gclm_debug_stamps.restype = ctypes.c_int
assert eng.lib.gclm_debug_stamps(eng._h, buf, n) == 0
allb = np.array(buf[:], dtype=np.int64)
st = allb[:(iters + 1) * ncta * 8].reshape(iters + 1, ncta, 8)
sol = allb[400000:400000 + (iters + 1) * 8].reshape(iters + 1, 8)
mhz = 1965.0
print(f"tiles {ncta} (rows_per_tile {cfg[2]}, warps {cfg[0]}, rsub {cfg[1]}), schedule {sched}")
names = ["prologue", "stream", "cta-reduce", "ticket", "sum partials", "solver+vote", "barrier"]
for p in range(iters + 1):
    s = st[p]
    last = int(np.argmax(s[:, 6]))  # the CTA that ran the solver has the latest stamp 6... all have; pick max stamp 4
    last = int(np.argmax(s[:, 4]))
    d = np.diff(s[last, :8].astype(np.float64)) / mhz
    others = (s[:, 3] - s[:, 0]).astype(np.float64) / mhz
    if p < iters:
        sd = np.diff(sol[p, :6].astype(np.float64)) / mhz
        print(f"         solver: cost+lambda+cholesky {sd[0]:.2f}, gravity {sd[1]:.2f}, focal/dist {sd[2]:.2f}, stores {sd[3]:.2f}, vote {sd[4]:.2f}")
    print(f"pass {p:2d}: " + ", ".join(f"{nm} {x:6.2f}" for nm, x in zip(names, d)) + f" | all CTAs start->reduce: median {np.median(others):.2f} max {others.max():.2f} us")
