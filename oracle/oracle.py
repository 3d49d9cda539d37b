"""ctypes front-end of the CPU oracle (oracle/gclm_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of gclm_oracle.c.  Only tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs may import this module.  The product package
(geocalib_b200/) must never import it.

Host-side pieces that the reference does in O(B) torch code are restated here in numpy:
  * trivial_init()  <- geocalib/lm_optimizer.py:20-58 (get_trivial_estimation),
                       geocalib/camera.py:49-93 (from_dict), geocalib/utils.py:272-279
  * gravity_from_rp / gravity_rp <- geocalib/gravity.py:29-40, 62-82
"""

from __future__ import annotations

import ctypes
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libgclm_oracle.so"
MODELS = {"pinhole": 0, "simple_radial": 1, "radial": 2, "simple_divisional": 3}
NUM_DIST = {"pinhole": 0, "simple_radial": 1, "radial": 2, "simple_divisional": 1}


def build(force: bool = False) -> Path:
    """Compile the oracle with oracle/Makefile (gcc only; a few seconds)."""
    src = _HERE / "gclm_oracle.c"
    if force or not _LIB_PATH.exists() or _LIB_PATH.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE), "-B"], check=True, capture_output=True)
    return _LIB_PATH


class _Conf(ctypes.Structure):
    _fields_ = [
        (n, ctypes.c_int)
        for n in (
            "model batch height width est_gravity est_focal apply_dist shared_intrinsics "
            "num_steps early_stop fix_lambda spherical log_focal global_semantics "
            "compute_uncertainty"
        ).split()
    ] + [(n, ctypes.c_double) for n in "lambda0 atol rtol up_scale lat_scale".split()]


_lib = None


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        build()
        try:
            _lib = ctypes.CDLL(str(_LIB_PATH))
        except OSError:
            build(force=True)
            _lib = ctypes.CDLL(str(_LIB_PATH))
    return _lib


def _ptr(a, dtype):
    if a is None:
        return None
    assert a.dtype == dtype and a.flags["C_CONTIGUOUS"], (a.dtype, a.flags)
    return a.ctypes.data_as(ctypes.c_void_p)


def _real(dtype):
    dtype = np.dtype(dtype)
    assert dtype in (np.float32, np.float64)
    return dtype, ("_f32" if dtype == np.float32 else "_f64")


DEFAULT_CONF = dict(  # geocalib/lm_optimizer.py:144-162
    camera_model="pinhole",
    shared_intrinsics=False,
    num_steps=30,
    lambda_=0.1,
    fix_lambda=False,
    early_stop=True,
    atol=1e-8,
    rtol=1e-8,
    use_spherical_manifold=True,
    use_log_focal=True,
    up_loss_fn_scale=1e-2,
    lat_loss_fn_scale=1e-2,
)


def system(model, cam, grav, up, lat, up_conf, lat_conf, *, rpf_basis=False, spherical=True,
           log_focal=True, up_scale=1e-2, lat_scale=1e-2, dtype=np.float32):
    """G (NF,), H (NF,NF), mean costs (2,) for ONE image with full columns [g0 g1 f k..]."""
    dtype, sfx = _real(dtype)
    nf = 3 + NUM_DIST[model]
    ref = up if up is not None else lat
    h, w = ref.shape[-2:]
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
    up, lat, up_conf, lat_conf, cam, grav = map(c, (up, lat, up_conf, lat_conf, cam, grav))
    G = np.zeros(nf, dtype)
    H = np.zeros((nf, nf), dtype)
    cost = np.zeros(2, dtype)
    fn = getattr(lib(), "gclm_oracle_system" + sfx)
    fn.restype = None
    fn(ctypes.c_int(MODELS[model]), ctypes.c_int(h), ctypes.c_int(w), _ptr(up, dtype),
       _ptr(lat, dtype), _ptr(up_conf, dtype), _ptr(lat_conf, dtype), ctypes.c_double(up_scale),
       ctypes.c_double(lat_scale), _ptr(cam, dtype), _ptr(grav, dtype), ctypes.c_int(rpf_basis),
       ctypes.c_int(spherical), ctypes.c_int(log_focal), _ptr(G, dtype), _ptr(H, dtype),
       _ptr(cost, dtype))
    return G, H, cost


def render(model, cam, grav, h, w, dtype=np.float32):
    """Noise-free up (2,h,w) / latitude (1,h,w) fields of one camera (get_perspective_field)."""
    dtype, sfx = _real(dtype)
    cam = np.ascontiguousarray(cam, dtype)
    grav = np.ascontiguousarray(grav, dtype)
    up = np.zeros((2, h, w), dtype)
    lat = np.zeros((1, h, w), dtype)
    fn = getattr(lib(), "gclm_oracle_render" + sfx)
    fn.restype = None
    fn(ctypes.c_int(MODELS[model]), ctypes.c_int(h), ctypes.c_int(w), _ptr(cam, dtype),
       _ptr(grav, dtype), _ptr(up, dtype), _ptr(lat, dtype))
    return up, lat


def gravity_from_rp(roll, pitch):
    roll, pitch = np.asarray(roll, np.float64), np.asarray(pitch, np.float64)
    sr, cr, sp, cp = np.sin(roll), np.cos(roll), np.sin(pitch), np.cos(pitch)
    g = np.stack([-sr * cp, -cr * cp, sp], -1)
    return g / np.maximum(np.linalg.norm(g, axis=-1, keepdims=True), 1e-12)


def gravity_rp(g):
    """roll/pitch read-back with the reference's 1e-4 eps (gravity.py:15,62-67)."""
    g = np.asarray(g)
    roll = np.arcsin(-g[..., 0] / (np.sqrt(1 - g[..., 2] ** 2) + 1e-4))
    offset = -np.pi * np.sign(g[..., 0])
    roll = np.where(g[..., 1] < 0, roll, -roll + offset)
    return np.stack([roll, np.arcsin(g[..., 2])], -1)


def gt_params(model, B, h, w, seed):
    """Ground-truth cameras (B,8) / gravities (B,3) of the synthetic workload: roll, pitch
    ~ U(-45, 45) deg, vfov ~ U(30, 90) deg (fisheye: U(60, 120)), k1 ~ U(-0.15, 0.15)
    (simple_divisional: U(-0.4, -0.05))."""
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


def procedural_fields(model, B, h, w, seed, sigma=0.02, conf_lo=0.2):
    """Seeded synthetic predicted fields, reproducible from (model, B, h, w, seed) alone: GT render in
    fp64 (rounded to fp32) + N(0, sigma^2) noise, up re-normalised, confidences ~ U(conf_lo, 1).

    Used by oracle/make_golden.py to feed the REFERENCE at the BASELINE image sizes (only the recipe and
    the reference's outputs are stored: the inputs would be 6-12 MB per case) and by the tests to
    regenerate the same inputs; `field_checksums` guards the regeneration.
    Returns (fields dict of float32 arrays, cam_gt (B,8), grav_gt (B,3))."""
    cam, grav = gt_params(model, B, h, w, seed)
    rng = np.random.default_rng(seed + 1)
    up = np.empty((B, 2, h, w), np.float32)
    lat = np.empty((B, 1, h, w), np.float32)
    for b in range(B):
        u, l = render(model, cam[b], grav[b], h, w, dtype=np.float64)
        u = u + rng.normal(0, sigma, (2, h, w))
        up[b] = (u / np.linalg.norm(u, axis=0, keepdims=True)).astype(np.float32)
        lat[b] = (l + rng.normal(0, sigma, (1, h, w))).astype(np.float32)
    uc = rng.uniform(conf_lo, 1, (B, h, w)).astype(np.float32)
    lc = rng.uniform(conf_lo, 1, (B, h, w)).astype(np.float32)
    return dict(up_field=up, latitude_field=lat, up_confidence=uc, latitude_confidence=lc), cam, grav


def field_checksums(fields):
    """Order-independent fp64 digests of the four field tensors (sum and sum of squares per tensor)."""
    out = []
    for k in ("up_field", "latitude_field", "up_confidence", "latitude_confidence"):
        a = fields[k].astype(np.float64)
        out += [a.sum(), (a * a).sum()]
    return np.array(out)


def trivial_init(batch, h, w, *, scales=None, prior_focal=None, prior_gravity=None,
                 prior_dist=None, dtype=np.float32):
    """Initial camera (B,8) / gravity (B,3): roll=pitch=0, f=0.7*max(h,w) via a vfov round trip."""
    f32 = np.float32
    hh = np.full(batch, h, f32)
    ww = np.full(batch, w, f32)
    focal = f32(0.7) * np.maximum(hh, ww) if prior_focal is None else np.asarray(prior_focal, f32)
    focal = np.broadcast_to(focal, (batch,)).astype(f32)
    vfov = (f32(2) * np.arctan(hh / (f32(2) * focal))).astype(f32)  # focal2fov
    f = (hh / f32(2) / np.tan(vfov / f32(2))).astype(f32)  # fov2focal
    fx = f.copy()
    if scales is not None:
        s = np.asarray(scales, f32)
        fx = (fx * s[..., 0] / s[..., 1]).astype(f32)
    k1 = np.zeros(batch, f32)
    k2 = np.zeros(batch, f32)
    if prior_dist is not None:
        d = np.asarray(prior_dist, f32).reshape(batch, -1)
        k1 = d[:, 0]
        k2 = d[:, 1] if d.shape[1] == 2 else k2
    cam = np.stack([ww, hh, fx, f, ww / f32(2), hh / f32(2), k1, k2], -1).astype(dtype)
    if prior_gravity is not None:
        g = np.asarray(prior_gravity, np.float64).reshape(batch, 3)
        g = (g / np.maximum(np.linalg.norm(g, axis=-1, keepdims=True), 1e-12)).astype(dtype)
    else:
        g = gravity_from_rp(np.zeros(batch, f32), np.zeros(batch, f32)).astype(dtype)
    return cam, g


def lm(model, up, lat, up_conf=None, lat_conf=None, *, conf=None, cam0=None, grav0=None,
       scales=None, prior_focal=None, prior_gravity=None, prior_dist=None, shared_intrinsics=None,
       global_semantics=True, compute_uncertainty=True, trace=False, dtype=np.float32,
       num_threads=None):
    """Full LM calibration of a batch on the CPU.  Mirrors LMOptimizer.forward (:646-664)."""
    dtype, sfx = _real(dtype)
    cf = {**DEFAULT_CONF, **(conf or {})}
    if shared_intrinsics is None:
        shared_intrinsics = cf["shared_intrinsics"]
    ref = up if up is not None else lat
    B, _, h, w = ref.shape
    c = lambda a: None if a is None else np.ascontiguousarray(a, dtype=dtype)
    up, lat, up_conf, lat_conf = map(c, (up, lat, up_conf, lat_conf))
    if cam0 is None:
        cam0, g0 = trivial_init(B, h, w, scales=scales, prior_focal=prior_focal,
                                prior_gravity=prior_gravity, prior_dist=prior_dist, dtype=dtype)
        grav0 = g0 if grav0 is None else grav0
    cam = np.ascontiguousarray(cam0, dtype).copy()
    grav = np.ascontiguousarray(grav0, dtype).copy()
    K = NUM_DIST[model]
    est_g, est_f = prior_gravity is None, prior_focal is None
    n = 2 * est_g + est_f + K
    cc = _Conf(
        model=MODELS[model], batch=B, height=h, width=w, est_gravity=est_g, est_focal=est_f,
        apply_dist=int(K > 0 and prior_dist is None), shared_intrinsics=int(shared_intrinsics),
        num_steps=cf["num_steps"], early_stop=int(cf["early_stop"]),
        fix_lambda=int(cf["fix_lambda"]), spherical=int(cf["use_spherical_manifold"]),
        log_focal=int(cf["use_log_focal"]), global_semantics=int(global_semantics),
        compute_uncertainty=int(compute_uncertainty), lambda0=cf["lambda_"], atol=cf["atol"],
        rtol=cf["rtol"], up_scale=cf["up_loss_fn_scale"], lat_scale=cf["lat_loss_fn_scale"],
    )
    costs = np.zeros((B, 6), dtype)
    stop_at = np.zeros(B, np.int32)
    status = np.zeros(B, np.int32)
    cov = np.zeros((B, n, n), dtype)
    unc = np.zeros((B, 5), dtype)
    tr = np.zeros((cf["num_steps"], B, 13), dtype) if trace else None
    fn = getattr(lib(), "gclm_oracle_lm" + sfx)
    fn.restype = ctypes.c_int
    old = os.environ.get("OMP_NUM_THREADS")
    if num_threads is not None:
        try:
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(num_threads))
        except OSError:
            pass
    iters = fn(ctypes.byref(cc), _ptr(up, dtype), _ptr(lat, dtype), _ptr(up_conf, dtype),
               _ptr(lat_conf, dtype), _ptr(cam, dtype), _ptr(grav, dtype), _ptr(costs, dtype),
               _ptr(stop_at, np.int32), _ptr(cov, dtype), _ptr(unc, dtype),
               _ptr(status, np.int32), _ptr(tr, dtype) if trace else None)
    del old
    out = dict(
        camera=cam, gravity=grav, covariance=cov, stop_at=stop_at, status=status, iters=iters,
        roll_uncertainty=unc[:, 0], pitch_uncertainty=unc[:, 1], gravity_uncertainty=unc[:, 2],
        focal_uncertainty=unc[:, 3], vfov_uncertainty=unc[:, 4],
        initial_cost=costs[:, 0], initial_up_cost=costs[:, 1], initial_latitude_cost=costs[:, 2],
        final_cost=costs[:, 3], final_up_cost=costs[:, 4], final_latitude_cost=costs[:, 5],
    )
    if trace:
        out["trace"] = tr[:iters]
    return out
