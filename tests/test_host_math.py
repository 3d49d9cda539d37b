"""The CUDA source's per-pixel math, compiled for the HOST, against the oracle -- no GPU needed.

tests/host_math/host_math.cu includes geocalib_b200/csrc/gclm_device.cuh and drives the same
`row_accumulate` the pass kernel's streaming loop calls (packed and scalar builds, log-focal and general focal
direction, the divisional k1 == 0 / clamped-root special cases).  Its sums over the pixels are compared with
the fp64 oracle system (G, H, costs) at an estimate that is NOT the one the fields were rendered from, so that
residuals, Huber weights (inliers and outliers) and every Jacobian column matter.
"""
import ctypes
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import oracle

HERE = Path(__file__).resolve().parent / "host_math"
MODELS = ["pinhole", "simple_radial", "radial", "simple_divisional"]
NUM_DIST = {"pinhole": 0, "simple_radial": 1, "radial": 2, "simple_divisional": 1}


@pytest.fixture(scope="module")
def hostmath():
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available")
    src, lib = HERE / "host_math.cu", HERE / "libgclm_hostmath.so"
    hdr = HERE.parents[1] / "geocalib_b200" / "csrc" / "gclm_device.cuh"
    if not lib.exists() or lib.stat().st_mtime < max(src.stat().st_mtime, hdr.stat().st_mtime):
        subprocess.run([nvcc, "-O2", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-Xcompiler", "-fPIC",
                        "-Xcompiler", "-Wno-unknown-pragmas", "-shared", "-o", str(lib), str(src)], check=True)
    L = ctypes.CDLL(str(lib))
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    L.gclm_host_math_system.restype = ctypes.c_int
    L.gclm_host_math_system.argtypes = [ctypes.c_int] * 3 + [fp] * 6 + [ctypes.c_int] * 5 + [ctypes.c_float] * 2 + [dp]
    L.gclm_host_math_adjoint.restype = ctypes.c_int
    L.gclm_host_math_adjoint.argtypes = [ctypes.c_int] * 3 + [fp] * 7 + [ctypes.c_float] * 2 + [fp]
    return L


def gt_and_estimate(model, h, w, seed, k1=None):
    rng = np.random.default_rng(seed)
    fish = model == "simple_divisional"
    def cam_of(vfov_deg, k1v, k2v):
        fy = h / 2 / np.tan(np.deg2rad(vfov_deg) / 2)
        return np.array([w, h, fy * 1.05, fy, w / 2, h / 2, k1v, k2v], np.float32)
    k1_gt = {"pinhole": 0.0, "simple_radial": 0.08, "radial": -0.1, "simple_divisional": -0.2}[model]
    k2_gt = 0.02 if model == "radial" else 0.0
    cam_gt = cam_of(95 if fish else 60, k1_gt, k2_gt)
    grav_gt = oracle.gravity_from_rp(0.3, -0.25).astype(np.float32)
    # the estimate the system is evaluated at: a few degrees / 15 % away from the truth
    k1_est = k1_gt * 0.5 if k1 is None else k1
    cam = cam_of(105 if fish else 70, k1_est, 0.5 * k2_gt)
    grav = oracle.gravity_from_rp(0.22, -0.15).astype(np.float32)
    up, lat = oracle.render(model, cam_gt, grav_gt, h, w)
    up = up + rng.normal(0, 0.02, up.shape).astype(np.float32)
    up /= np.linalg.norm(up, axis=0, keepdims=True)
    lat = lat + rng.normal(0, 0.02, lat.shape).astype(np.float32)
    # a block of gross outliers: Huber's linear branch
    up[:, : h // 4, : w // 4] = np.array([0.6, -0.8], np.float32)[:, None, None]
    lat[:, : h // 4, w // 2:] += 0.5
    uconf = rng.uniform(0.2, 1.0, (h, w)).astype(np.float32)
    lconf = rng.uniform(0.2, 1.0, (h, w)).astype(np.float32)
    return cam, grav, np.ascontiguousarray(up, np.float32), np.ascontiguousarray(lat, np.float32), uconf, lconf


def host_system(L, model, cam, grav, up, lat, uconf, lconf, *, spherical, log_focal, rpf, pack, force_gen_d=False):
    h, w = lat.shape[-2:]
    nf = 3 + NUM_DIST[model]
    nh = nf * (nf + 1) // 2
    out = np.zeros(nh + nf + 2, np.float64)
    fp = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    n = L.gclm_host_math_system(MODELS.index(model), h, w, fp(up), fp(lat), fp(uconf), fp(lconf), fp(cam), fp(grav),
                                int(spherical), int(log_focal), int(rpf), int(pack), int(force_gen_d), 1e-2, 1e-2,
                                out.ctypes.data_as(ctypes.POINTER(ctypes.c_double)))
    assert n == out.size
    H = np.zeros((nf, nf))
    H[np.triu_indices(nf)] = out[:nh]
    H = H + np.triu(H, 1).T
    costs = out[nh + nf:] * 1e-4 / (h * w)  # the epilogue's scale^2 / N
    return out[nh:nh + nf], H, costs


def assert_system_close(got, ref, tol=3e-6):
    """fp32 pixel math vs the fp64 oracle: measured differences are ~1e-7 in the scale of the system."""
    G, H, c = got
    Gr, Hr, cr = (np.asarray(a, np.float64) for a in ref)
    # compare in the scale of the system (sqrt of the diagonal), as the GPU parity test does
    d = np.sqrt(np.maximum(np.diag(Hr), 1e-30))
    assert np.abs((H - Hr) / np.outer(d, d)).max() < tol, np.abs((H - Hr) / np.outer(d, d)).max()
    assert np.abs((G - Gr) / d).max() < tol * np.abs(Gr / d).max() + 1e-6, np.abs((G - Gr) / d).max()
    assert np.allclose(c, cr, rtol=2e-6, atol=1e-12), (c, cr)


@pytest.mark.parametrize("pack", [1, 0], ids=["packed", "scalar"])
@pytest.mark.parametrize("basis", ["lm_pass", "linear_focal_euclidean", "rpf_final"])
@pytest.mark.parametrize("model", MODELS)
def test_kernel_math_matches_oracle(hostmath, model, basis, pack):
    h, w = 48, 64
    cam, grav, up, lat, uconf, lconf = gt_and_estimate(model, h, w, 3)
    kw = {"lm_pass": dict(spherical=True, log_focal=True, rpf=False),
          "linear_focal_euclidean": dict(spherical=False, log_focal=False, rpf=False),
          "rpf_final": dict(spherical=True, log_focal=True, rpf=True)}[basis]
    got = host_system(hostmath, model, cam, grav, up, lat, uconf, lconf, pack=pack, **kw)
    ref = oracle.system(model, cam, grav, up, lat[None], uconf, lconf, rpf_basis=kw["rpf"], spherical=kw["spherical"],
                        log_focal=kw["log_focal"], dtype=np.float64)
    assert_system_close(got, ref)
    if basis == "lm_pass":  # the general-focal-direction code must agree with the log-focal specialisation
        gen = host_system(hostmath, model, cam, grav, up, lat, uconf, lconf, pack=pack, force_gen_d=True, **kw)
        assert_system_close(gen, ref)


@pytest.mark.parametrize("pack", [1, 0], ids=["packed", "scalar"])
@pytest.mark.parametrize("rpf", [False, True], ids=["lm_pass", "rpf_final"])
def test_divisional_k1_zero_masked_branches(hostmath, pack, rpf):
    """k1 == 0 (trivial init / zero prior): the reference's masked denominators; the packed build
    uses the constant form (model_scalars<..., K0>), the scalar build the literal formulas; centre pixel r^2 = 0."""
    h, w = 48, 64  # even: pixel (w/2, h/2) has u = v = 0
    cam, grav, up, lat, uconf, lconf = gt_and_estimate("simple_divisional", h, w, 5, k1=0.0)
    got = host_system(hostmath, "simple_divisional", cam, grav, up, lat, uconf, lconf, spherical=True, log_focal=True,
                      rpf=rpf, pack=pack)
    ref = oracle.system("simple_divisional", cam, grav, up, lat[None], uconf, lconf, rpf_basis=rpf, dtype=np.float64)
    assert_system_close(got, ref)


@pytest.mark.parametrize("pack", [1, 0], ids=["packed", "scalar"])
def test_divisional_clamped_root_pixels(hostmath, pack):
    """k1 > 0 large enough that 1 - 4 k1 r^2 < 1e-6 in the image corners: those pixels take the literal path."""
    h, w = 48, 64
    cam, grav, up, lat, uconf, lconf = gt_and_estimate("simple_divisional", h, w, 7, k1=0.45)
    fy = cam[3]
    r2_corner = (w / 2 / cam[2]) ** 2 + (h / 2 / fy) ** 2
    assert 1 - 4 * 0.45 * r2_corner < 0, "the test image must contain clamped pixels"
    got = host_system(hostmath, "simple_divisional", cam, grav, up, lat, uconf, lconf, spherical=True, log_focal=True,
                      rpf=False, pack=pack)
    ref = oracle.system("simple_divisional", cam, grav, up, lat[None], uconf, lconf, dtype=np.float64)
    assert_system_close(got, ref, tol=2e-4)  # Jacobians blow up next to the clamp (condition number ~1e10)


ADJOINT_CASES = [(m, None) for m in MODELS] + [("simple_divisional", 0.0), ("simple_divisional", 0.45)]


@pytest.mark.parametrize("scale", [1e-2, 0.2], ids=["default_scale_mostly_outliers", "all_inliers"])
@pytest.mark.parametrize("model,k1", ADJOINT_CASES, ids=[f"{m}{'' if k is None else '_k1=' + str(k)}" for m, k in ADJOINT_CASES])
def test_adjoint_pixel_is_the_derivative_of_the_oracle_gradient(hostmath, model, k1, scale):
    """The per-pixel body of the differentiable LM's adjoint kernel (`adjoint_pixel`, host-compiled) against fp64
    central differences of lambda . G of the ORACLE system, element by element over all five input planes: Jacobian rows,
    Huber inlier and outlier branches (dw/do), confidences; for simple_divisional also the literal masked formulas
    (k1 == 0: every pixel; k1 = 0.45: the clamped-root pixels in the corners)."""
    h, w = 12, 16
    cam, grav, up, lat, uconf, lconf = gt_and_estimate(model, h, w, 11, k1=k1)
    nf = 3 + NUM_DIST[model]
    lam = np.random.default_rng(2).normal(size=nf).astype(np.float32)
    out = np.zeros((5, h, w), np.float32)
    fp = lambda a: np.ascontiguousarray(a, np.float32).ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    assert hostmath.gclm_host_math_adjoint(MODELS.index(model), h, w, fp(up), fp(lat), fp(uconf), fp(lconf), fp(cam), fp(grav),
                                           fp(lam), scale, scale, out.ctypes.data_as(ctypes.POINTER(ctypes.c_float))) == 0
    planes = [up.astype(np.float64), lat.astype(np.float64), uconf.astype(np.float64), lconf.astype(np.float64)]

    def lam_G(pl):
        G, _, _ = oracle.system(model, cam, grav, pl[0], pl[1][None], pl[2], pl[3], up_scale=scale, lat_scale=scale, dtype=np.float64)
        return float(np.dot(lam.astype(np.float64), G))

    eps = 1e-6
    fd = np.zeros((5, h, w))
    for k, (pi, ch) in enumerate([(0, 0), (0, 1), (1, None), (2, None), (3, None)]):
        for y in range(h):
            for x in range(w):
                idx = (ch, y, x) if ch is not None else (..., y, x)
                vals = []
                for sgn in (+1, -1):
                    pl = [p.copy() for p in planes]
                    pl[pi][idx] += sgn * eps
                    vals.append(lam_G(pl))
                fd[k, y, x] = (vals[0] - vals[1]) / (2 * eps)
    for k, name in enumerate(["up_x", "up_y", "latitude", "up_confidence", "latitude_confidence"]):
        err = np.abs(out[k] - fd[k]).max() / max(np.abs(fd[k]).max(), 1e-12)
        # fp32 pixel math: measured <= 3e-5 of the plane maximum; next to the clamped root (k1 = 0.45: most pixels of this
        # 12x16 fisheye) the formulas amplify rounding by ~1e3 (cf. test_divisional_clamped_root_pixels): 3e-4 measured
        assert err < (1e-3 if k1 == 0.45 else 1e-4), (name, err, np.abs(fd[k]).max())
    if scale < 0.1:  # both Huber branches are exercised (latitude outliers have a zero derivative: w r = conf a sign(r))
        assert (np.abs(fd[2]) == 0).any() and (np.abs(fd[2]) > 0).any()


def _bind_solver(L):
    fp = ctypes.POINTER(ctypes.c_float)
    L.gclm_host_math_solve.restype = ctypes.c_int
    L.gclm_host_math_solve.argtypes = [ctypes.c_int, fp, fp, fp]
    L.gclm_host_math_spd_inverse.restype = ctypes.c_int
    L.gclm_host_math_spd_inverse.argtypes = [ctypes.c_int, fp, fp]
    L.gclm_host_math_update.restype = None
    L.gclm_host_math_update.argtypes = [ctypes.c_int, fp, fp, fp, ctypes.c_int, ctypes.c_int]
    return fp


@pytest.mark.parametrize("n", [3, 4, 5])
def test_fixed_size_cholesky_and_spd_inverse(hostmath, n):
    """`cholesky_solve_fixed<N>` (the LM step) and `spd_inverse_fixed<N>` (the covariance) against numpy in fp64, on
    systems with the scaling of real ones (gravity ~1e4, focal ~1e3, distortion ~1e2..1e4) and the LM damping."""
    fp = _bind_solver(hostmath)
    rng = np.random.default_rng(n)
    for trial in range(20):
        J = rng.normal(size=(200, n)) * np.array([100.0, 90.0, 30.0, 50.0, 150.0])[:n]
        H = J.T @ J
        H += np.diag(np.maximum(0.1 * np.diag(H), 1e-6))   # lm_optimizer.py:116-121
        b = rng.normal(size=n) * np.sqrt(np.diag(H))
        A32, b32 = np.ascontiguousarray(H, np.float32), np.ascontiguousarray(b, np.float32)
        x, Ai = np.zeros(n, np.float32), np.zeros((n, n), np.float32)
        assert hostmath.gclm_host_math_solve(n, A32.ctypes.data_as(fp), b32.ctypes.data_as(fp), x.ctypes.data_as(fp)) == 0
        assert hostmath.gclm_host_math_spd_inverse(n, A32.ctypes.data_as(fp), Ai.ctypes.data_as(fp)) == 0
        xr = np.linalg.solve(A32.astype(np.float64), b32.astype(np.float64))
        Air = np.linalg.inv(A32.astype(np.float64))
        cond = np.linalg.cond(A32.astype(np.float64))
        assert np.abs(x - xr).max() <= 4e-7 * cond * np.abs(xr).max(), (trial, cond)
        d = np.sqrt(np.diag(Air))
        assert np.abs((Ai - Air) / np.outer(d, d)).max() <= 4e-7 * cond, (trial, cond)
    # not positive definite: reported, not NaN
    bad = np.diag([1.0, -1.0, 1.0, 1.0, 1.0][:n]).astype(np.float32)
    out = np.zeros(n, np.float32)
    assert hostmath.gclm_host_math_solve(n, bad.ctypes.data_as(fp), out.ctypes.data_as(fp), out.ctypes.data_as(fp)) == 1


@pytest.mark.parametrize("model", MODELS)
def test_parameter_updates_against_the_oracle_conventions(hostmath, model):
    """gravity_update / focal_update / dist_update (lm_optimizer.py:518-549): sphere retraction, log-focal step with the
    vfov clamp at [5, 150] degrees, distortion clamp, delta k1 mirrored into the k2 slot of one-parameter models."""
    import torch

    from geocalib_b200 import autograd as AG

    fp = _bind_solver(hostmath)
    K = NUM_DIST[model]
    rng = np.random.default_rng(5)
    for trial in range(50):
        g = rng.normal(size=3)
        g /= np.linalg.norm(g)
        h = 480.0
        cam = np.array([640, h, 500.0, 480.0, 320, 240, 0.1 if K else 0.0, -0.02 if K == 2 else (0.1 if K == 1 else 0.0)], np.float32)
        delta = np.concatenate([rng.normal(size=2) * 0.2, rng.normal(size=1) * (3.0 if trial % 5 == 0 else 0.3),
                                rng.normal(size=K) * (2.0 if trial % 7 == 0 else 0.1)]).astype(np.float32)
        cam_new, g_new = cam.copy(), g.astype(np.float32).copy()
        hostmath.gclm_host_math_update(MODELS.index(model), cam_new.ctypes.data_as(fp), g_new.ctypes.data_as(fp),
                                       np.ascontiguousarray(np.concatenate([delta, np.zeros(2, np.float32)])).ctypes.data_as(fp), 1, 1)
        g_ref = AG.sphere_plus(torch.from_numpy(g.astype(np.float32))[None], torch.from_numpy(delta[None, :2]))[0].numpy()
        assert np.abs(g_new - g_ref).max() < 2e-6
        fy = np.clip(np.exp(np.log(480.0) + float(delta[2])), h / 2 / np.tan(np.deg2rad(75.0)), h / 2 / np.tan(np.deg2rad(2.5)))
        assert abs(cam_new[3] - fy) <= 2e-6 * fy and abs(cam_new[2] - fy * 500.0 / 480.0) <= 3e-6 * fy
        lim = 3.0 if model == "simple_divisional" else 0.7
        if K >= 1:
            assert abs(cam_new[6] - np.clip(cam[6] + delta[3], -lim, lim)) < 1e-6
        if K == 1:
            assert cam_new[7] == cam_new[6]
        if K == 2:
            assert abs(cam_new[7] - np.clip(cam[7] + delta[4], -lim, lim)) < 1e-6
        if K == 0:
            assert cam_new[6] == 0 and cam_new[7] == 0
        assert np.array_equal(cam_new[[0, 1, 4, 5]], cam[[0, 1, 4, 5]])
