"""Generate golden vectors for the LM calibration path by RUNNING THE REFERENCE.

Run in the build container only (the reference lives at /root/reference, read-only, and does not
travel to the GPU box):

    python oracle/make_golden.py            # writes tests/golden/*.npz

The reference is imported unmodified; the only shim is an empty `kornia` module because
geocalib/utils.py imports kornia at module scope (it is only *used* by ImagePreprocessor.resize,
which is not on this path).  Everything stored here is an input or an output of the reference's
own functions:

  system_<model>.npz   LMOptimizer.setup_system / calculate_costs at fixed parameters
                       (fp32 and fp64; optimisation basis and as_rpf basis)
  lm_<case>.npz        LMOptimizer(conf).eval().forward(data) end to end
  init.npz             get_trivial_estimation for a few shapes / priors / scales
  base_<case>.npz      BASELINE.json configs at their real image sizes (640x480, 1024x1024, and the
                       single 320x240 image of config 0): the inputs are NOT stored (6-12 MB each) but
                       regenerated from a seeded recipe (oracle.procedural_fields, checksummed);
                       stored are the reference's outputs, its iteration-0 system (G, H, costs in
                       fp64 at the trivial initial estimate) and 1/2/3-iteration trajectories
                       (`python oracle/make_golden.py --baseline` writes only these)

TEST INFRASTRUCTURE ONLY.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

import numpy as np
import torch

sys.modules.setdefault("kornia", types.ModuleType("kornia"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))

from geocalib.camera import camera_models  # noqa: E402
from geocalib.gravity import Gravity  # noqa: E402
from geocalib.lm_optimizer import LMOptimizer, get_trivial_estimation  # noqa: E402
from geocalib.perspective_fields import get_perspective_field  # noqa: E402

OUT = Path(__file__).resolve().parent.parent / "tests" / "golden"
MODELS = ["pinhole", "simple_radial", "radial", "simple_divisional"]


def make_params(model, B, h, w, rng, fisheye=False, fxfy=True):
    roll = rng.uniform(-0.6, 0.6, B)
    pitch = rng.uniform(-0.6, 0.6, B)
    vfov = np.deg2rad(rng.uniform(60, 110, B) if fisheye else rng.uniform(35, 85, B))
    fy = h / 2 / np.tan(vfov / 2)
    fx = fy * (rng.uniform(0.9, 1.1, B) if fxfy else 1.0)
    k1 = np.zeros(B)
    k2 = np.zeros(B)
    if model in ("simple_radial", "radial"):
        k1 = rng.uniform(-0.15, 0.15, B)
    if model == "radial":
        k2 = rng.uniform(-0.03, 0.03, B)
    if model == "simple_divisional":
        k1 = rng.uniform(-0.35, -0.05, B)
    cam = np.stack([np.full(B, w), np.full(B, h), fx, fy, np.full(B, w / 2), np.full(B, h / 2), k1, k2], -1)
    return cam.astype(np.float32), roll.astype(np.float32), pitch.astype(np.float32)


def render_noisy(model, cam, roll, pitch, rng, sigma=0.02):
    camera = camera_models[model](torch.from_numpy(cam))
    gravity = Gravity.from_rp(torch.from_numpy(roll), torch.from_numpy(pitch)).float()
    up, lat = get_perspective_field(camera, gravity)
    up_gt, lat_gt = up.numpy().copy(), lat.numpy().copy()
    up = up_gt + rng.normal(0, sigma, up_gt.shape).astype(np.float32)
    up = up / np.linalg.norm(up, axis=1, keepdims=True)
    lat = lat_gt + rng.normal(0, sigma, lat_gt.shape).astype(np.float32)
    B, _, h, w = up.shape
    up_conf = rng.uniform(0.2, 1.0, (B, h, w)).astype(np.float32)
    lat_conf = rng.uniform(0.2, 1.0, (B, h, w)).astype(np.float32)
    return dict(up_field=up.astype(np.float32), latitude_field=lat.astype(np.float32),
                up_confidence=up_conf, latitude_confidence=lat_conf,
                up_gt=up_gt, lat_gt=lat_gt, gravity_gt=gravity.vec3d.numpy())


def t(a, double=False):
    x = torch.from_numpy(np.ascontiguousarray(a))
    return x.double() if double else x


def golden_system(model, seed):
    rng = np.random.default_rng(seed)
    h, w = 24, 32
    B = 3
    cam, roll, pitch = make_params(model, B, h, w, rng, fisheye=(model == "simple_divisional"))
    if model == "simple_divisional":
        cam[1, 6] = 0.0  # the trivial-init k1 == 0 branch (masked denominators)
        cam[2, 6] = 0.12
    f = render_noisy(model, cam, roll, pitch, rng)
    # evaluate the system away from the GT so residuals are O(0.1)
    cam_eval = cam.copy()
    cam_eval[:, 2:4] *= rng.uniform(0.8, 1.25, (B, 1)).astype(np.float32)
    if model != "pinhole" :
        cam_eval[:, 6] += np.where(cam[:, 6] == 0, 0, rng.uniform(-0.03, 0.03, B)).astype(np.float32)
    g_eval = Gravity.from_rp(t(roll + 0.1), t(pitch - 0.08)).float().vec3d.numpy()
    out = dict(model=model, cam=cam_eval, grav=g_eval,
               **{k: f[k] for k in ("up_field", "latitude_field", "up_confidence", "latitude_confidence")})
    opt = LMOptimizer({"camera_model": model}).eval()
    for tag, dbl in (("f32", False), ("f64", True)):
        camera = camera_models[model](t(cam_eval, dbl))
        gravity = Gravity(t(g_eval, dbl))
        data = {k: t(f[k], dbl) for k in ("up_field", "latitude_field", "up_confidence", "latitude_confidence")}
        res = opt.calculate_residuals(camera, gravity, data)
        costs, weights = opt.calculate_costs(res, data)
        for basis, as_rpf in (("opt", False), ("rpf", True)):
            G, H = opt.setup_system(camera, gravity, res, weights, as_rpf=as_rpf)
            out[f"G_{basis}_{tag}"] = G.numpy()
            out[f"H_{basis}_{tag}"] = H.numpy()
        out[f"cost_up_{tag}"] = costs["up_cost"].mean(-1).numpy()
        out[f"cost_lat_{tag}"] = costs["latitude_cost"].mean(-1).numpy()
    # noise-free render check for the oracle's render()
    out["cam_gt"], out["grav_gt"] = cam, f["gravity_gt"]
    out["up_gt"], out["lat_gt"] = f["up_gt"], f["lat_gt"]
    np.savez_compressed(OUT / f"system_{model}.npz", **out)


def run_reference(model, data_np, conf=None, shared=False, extra=None):
    conf = {"camera_model": model, **(conf or {})}
    opt = LMOptimizer(conf).eval()
    opt.shared_intrinsics = shared  # as extractor.py:113 does
    data = {k: t(v) for k, v in data_np.items()}
    data |= extra or {}
    with torch.no_grad():
        out = opt(data)
    res = {}
    for k, v in out.items():
        if k == "camera":
            res["camera"] = v._data.numpy()
        elif k == "gravity":
            res["gravity"] = v.vec3d.numpy()
            res["rp"] = v.rp.numpy()
        elif isinstance(v, torch.Tensor):
            res[k] = v.numpy()
        else:
            res[k] = np.asarray(v)
    return res


def golden_lm(name, model, seed, B, h, w, conf=None, shared=False, fields=None, use_conf=True,
              scales=None, prior_focal=False, prior_gravity=False, prior_dist=False, iters=(1, 2, 3),
              shared_f=False):
    rng = np.random.default_rng(seed)
    cam, roll, pitch = make_params(model, B, h, w, rng, fisheye=(model == "simple_divisional"),
                                   fxfy=False)
    if shared or shared_f:
        cam[:, 2:4] = cam[0, 2:4]
        cam[:, 6:] = cam[0, 6:]
    if scales is not None:
        cam[:, 2] = cam[:, 3] * scales[0] / scales[1]
    f = render_noisy(model, cam, roll, pitch, rng)
    keys = ["up_field", "latitude_field"] if fields is None else list(fields)
    if use_conf:
        keys += [k for k in ("up_confidence", "latitude_confidence") if k.split("_")[0] in "".join(keys)]
    data_np = {k: f[k] for k in keys}
    extra = {}
    save = dict(model=model, conf_keys=np.array(list((conf or {}).keys())),
                conf_vals=np.array([float(v) for v in (conf or {}).values()]),
                shared=shared, cam_gt=cam, grav_gt=f["gravity_gt"], **{f"in_{k}": v for k, v in data_np.items()})
    if scales is not None:
        extra["scales"] = t(np.asarray(scales, np.float32))
        save["scales"] = np.asarray(scales, np.float32)
    if prior_focal:
        pf = (cam[:, 3] * 1.03).astype(np.float32)
        extra["prior_focal"] = t(pf)
        save["prior_focal"] = pf
    if prior_gravity:
        pg = Gravity.from_rp(t(roll + 0.01), t(pitch - 0.01)).float().vec3d.numpy()
        extra["prior_gravity"] = t(pg)
        save["prior_gravity"] = pg
    if prior_dist:
        pd = np.array([cam[0, 6] * 0.9], np.float32)  # (K,) only works for B == 1 (camera.py:73-79)
        extra["prior_dist"] = t(pd)
        save["prior_dist"] = pd
    res = run_reference(model, data_np, conf, shared, extra)
    save |= {f"out_{k}": v for k, v in res.items()}
    # parameter trajectory: re-run with num_steps = k, early_stop off (deterministic reference)
    for k in iters:
        r = run_reference(model, data_np, {**(conf or {}), "num_steps": k, "early_stop": False}, shared, extra)
        save[f"it{k}_camera"] = r["camera"]
        save[f"it{k}_gravity"] = r["gravity"]
        save[f"it{k}_final_cost"] = r["final_cost"]
    np.savez_compressed(OUT / f"lm_{name}.npz", **save)
    print(name, "stop_at", res["stop_at"], "f", res["camera"][:, 3], "gt", cam[:, 3])


def golden_init():
    out = {}
    cases = [
        ("plain", 240, 320, {}),
        ("tall", 480, 320, {}),
        ("scales", 320, 320, {"scales": torch.tensor([0.5, 0.4])}),
        ("pfocal", 240, 320, {"prior_focal": torch.tensor([300.0, 200.0])}),
        ("pgrav", 240, 320, {"prior_gravity": torch.tensor([[0.1, -0.9, 0.2], [0.0, -1.0, 0.3]])}),
    ]
    for name, h, w, extra in cases:
        data = {"latitude_field": torch.zeros(2, 1, h, w), **extra}
        cam, grav = get_trivial_estimation(data, camera_models["pinhole"])
        out[f"{name}_cam"] = cam._data.numpy()
        out[f"{name}_grav"] = grav.vec3d.numpy()
        out[f"{name}_hw"] = np.array([h, w])
        for k, v in extra.items():
            out[f"{name}_{k}"] = v.numpy()
    data = {"latitude_field": torch.zeros(1, 1, 240, 320), "prior_dist": torch.tensor([0.07])}
    cam, _ = get_trivial_estimation(data, camera_models["simple_radial"])
    out["pdist_cam"] = cam._data.numpy()
    out["pdist_prior_dist"] = np.array([0.07], np.float32)
    out["pdist_hw"] = np.array([240, 320])
    np.savez_compressed(OUT / "init.npz", **out)


def golden_baseline(name, model, B, h, w, seed, conf, iters=(1, 2, 3)):
    """A BASELINE.json configuration at its real image size: recipe in, reference outputs out."""
    from oracle import oracle  # the seeded input recipe only (render + numpy noise)

    f, cam_gt, grav_gt = oracle.procedural_fields(model, B, h, w, seed)
    save = dict(model=model, B=B, h=h, w=w, seed=seed, conf_keys=np.array(list(conf.keys())),
                conf_vals=np.array([float(v) for v in conf.values()]), cam_gt=cam_gt, grav_gt=grav_gt,
                checksums=oracle.field_checksums(f))
    res = run_reference(model, f, conf)
    save |= {f"out_{k}": v for k, v in res.items()}
    for k in iters:
        r = run_reference(model, f, {**conf, "num_steps": k, "early_stop": False})
        save[f"it{k}_camera"] = r["camera"]
        save[f"it{k}_gravity"] = r["gravity"]
        save[f"it{k}_final_cost"] = r["final_cost"]
    # iteration-0 system at the trivial initial estimate, fp64 (SURVEY 8d "Parity check")
    opt = LMOptimizer({"camera_model": model, **conf}).eval()
    data32 = {k: t(v) for k, v in f.items()}
    cam0, grav0 = get_trivial_estimation(data32, camera_models[model])
    save["init_camera"], save["init_gravity"] = cam0._data.numpy(), grav0.vec3d.numpy()
    data = {k: t(v, True) for k, v in f.items()}
    camera, gravity = camera_models[model](cam0._data.double()), Gravity(grav0.vec3d.double())
    resid = opt.calculate_residuals(camera, gravity, data)
    costs, weights = opt.calculate_costs(resid, data)
    for basis, as_rpf in (("opt", False), ("rpf", True)):
        G, H = opt.setup_system(camera, gravity, resid, weights, as_rpf=as_rpf)
        save[f"G0_{basis}_f64"], save[f"H0_{basis}_f64"] = G.numpy(), H.numpy()
    save["cost0_up_f64"] = costs["up_cost"].mean(-1).numpy()
    save["cost0_lat_f64"] = costs["latitude_cost"].mean(-1).numpy()
    np.savez_compressed(OUT / f"base_{name}.npz", **save)
    print(name, "stop_at", res["stop_at"], "f", res["camera"][:, 3], "gt", cam_gt[:, 3])


def baseline_cases():
    fixed20 = {"num_steps": 20, "early_stop": False}
    golden_baseline("config0_pinhole_320x240", "pinhole", 1, 240, 320, 7000, {"num_steps": 10, "early_stop": False})
    golden_baseline("config1_pinhole_640x480", "pinhole", 4, 480, 640, 7001, fixed20)
    golden_baseline("config1_pinhole_640x480_default", "pinhole", 2, 480, 640, 7002, {})
    golden_baseline("config2_simple_radial_640x480", "simple_radial", 4, 480, 640, 7003, fixed20)
    golden_baseline("config3_simple_divisional_1024", "simple_divisional", 2, 1024, 1024, 7004, fixed20)
    golden_baseline("radial_640x480", "radial", 2, 480, 640, 7005, fixed20)


def main():
    OUT.mkdir(parents=True, exist_ok=True)
    torch.manual_seed(0)
    if "--baseline" in sys.argv:
        baseline_cases()
        return
    for i, m in enumerate(MODELS):
        golden_system(m, 100 + i)
    for i, m in enumerate(MODELS):
        golden_lm(f"default_{m}", m, 200 + i, B=3, h=48, w=64)
    golden_lm("fixed10_pinhole", "pinhole", 300, B=2, h=32, w=48, conf={"num_steps": 10, "early_stop": False})
    golden_lm("fixlambda_pinhole", "pinhole", 301, B=2, h=32, w=48, conf={"fix_lambda": True})
    golden_lm("euclid_pinhole", "pinhole", 302, B=2, h=32, w=48, conf={"use_spherical_manifold": False})
    golden_lm("linfocal_simple_radial", "simple_radial", 303, B=2, h=32, w=48, conf={"use_log_focal": False})
    golden_lm("scales_pinhole", "pinhole", 304, B=2, h=32, w=48, scales=(0.5, 0.4))
    golden_lm("priorfocal_pinhole", "pinhole", 305, B=2, h=32, w=48, prior_focal=True)
    golden_lm("priorgravity_pinhole", "pinhole", 306, B=2, h=32, w=48, prior_gravity=True)
    golden_lm("priorgravity_radial", "radial", 307, B=2, h=32, w=48, prior_gravity=True)
    golden_lm("priordist_simple_radial", "simple_radial", 308, B=1, h=32, w=48, prior_dist=True)
    golden_lm("latonly_pinhole", "pinhole", 310, B=2, h=32, w=48, fields=["latitude_field"])
    golden_lm("noconf_simple_divisional", "simple_divisional", 311, B=2, h=32, w=48, use_conf=False)
    golden_lm("shared_pinhole", "pinhole", 400, B=4, h=32, w=48, shared=True)
    golden_lm("shared_radial", "radial", 401, B=3, h=32, w=48, shared=True)
    golden_lm("shared_simple_divisional", "simple_divisional", 402, B=3, h=32, w=48, shared=True)
    golden_init()
    baseline_cases()


if __name__ == "__main__":
    main()
