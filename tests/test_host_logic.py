"""Host-side mirror of the reference interface: initial estimate, priors bookkeeping, Camera /
Gravity containers.  CPU only; the init is pinned against reference golden vectors."""
import numpy as np
import pytest
import torch

import geocalib_b200 as G
from oracle import oracle


def test_trivial_init_matches_reference(golden_dir):
    g = dict(np.load(golden_dir / "init.npz"))
    for name in ("plain", "tall", "scales", "pfocal", "pgrav"):
        h, w = g[name + "_hw"]
        data = {"latitude_field": torch.zeros(2, 1, h, w)}
        for k in ("scales", "prior_focal", "prior_gravity"):
            if f"{name}_{k}" in g:
                data[k] = torch.from_numpy(g[f"{name}_{k}"])
        cam, grav = G.get_trivial_estimation(data, G.Pinhole)
        np.testing.assert_array_equal(cam.numpy(), g[name + "_cam"])
        np.testing.assert_array_equal(grav.numpy(), g[name + "_grav"])
        # the oracle's numpy restatement of the same init
        kw = {k: g[f"{name}_{k}"] for k in ("scales", "prior_focal", "prior_gravity") if f"{name}_{k}" in g}
        c2, g2 = oracle.trivial_init(2, int(h), int(w), **kw)
        np.testing.assert_allclose(c2, g[name + "_cam"], rtol=2e-7)
        np.testing.assert_allclose(g2, g[name + "_grav"], atol=1e-7)
    data = {"latitude_field": torch.zeros(1, 1, 240, 320), "prior_dist": torch.from_numpy(g["pdist_prior_dist"])}
    cam, _ = G.get_trivial_estimation(data, G.SimpleRadial)
    np.testing.assert_array_equal(cam.numpy(), g["pdist_cam"])


def test_optimizer_surface_matches_reference():
    ref_keys = {"camera_model", "shared_intrinsics", "num_steps", "lambda_", "fix_lambda", "early_stop", "atol",
                "rtol", "use_spherical_manifold", "use_log_focal", "up_loss_fn_scale", "lat_loss_fn_scale", "verbose"}
    assert ref_keys <= set(G.LMOptimizer.default_conf)
    o = G.LMOptimizer({})
    assert o.num_steps == 30 and o.camera_model is G.Pinhole and not o.camera_has_distortion
    assert (o.gravity_delta_dims, o.focal_delta_dims, o.dist_delta_dims, o.n_intrinsic_params) == ((0, 1), (2,), None, 1)
    assert len(list(o.parameters())) == 0 and len(o.state_dict()) == 0  # no weights: state_dict unaffected
    o.set_camera_model("radial")
    o.setup_optimization_and_priors({}, shared_intrinsics=True)
    assert o.shared_intrinsics and o.camera_has_distortion
    assert (o.dist_delta_dims, o.n_intrinsic_params) == ((3, 4), 3)
    o.setup_optimization_and_priors({"prior_gravity": 1})
    assert (o.gravity_delta_dims, o.focal_delta_dims, o.dist_delta_dims) == ((-1,), (0,), (1, 2))
    o.setup_optimization_and_priors({"prior_focal": 1, "prior_dist": 1})
    assert not o.estimate_focal and not o.estimate_dist and o.n_intrinsic_params == 2
    with pytest.raises(AssertionError):
        o.set_camera_model("nope")


def test_cpu_tensors_fail_loudly():
    from geocalib_b200._lib import GclmError

    o = G.LMOptimizer({}).eval()
    with pytest.raises(GclmError, match="CUDA"):
        o({"up_field": torch.zeros(1, 2, 32, 32), "latitude_field": torch.zeros(1, 1, 32, 32)})
    with pytest.raises(KeyError):  # the reference needs latitude_field too (lm_optimizer.py:31)
        o({})


def test_camera_container():
    cam = G.Radial(torch.tensor([[640.0, 480.0, 500.0, 510.0, 320.0, 240.0, -0.1, 0.02]]))
    assert cam.shape == (1,) and cam.name() == "radial" and cam.num_dist_params() == 2
    assert torch.equal(cam.f, torch.tensor([[500.0, 510.0]])) and torch.equal(cam.dist, torch.tensor([[-0.1, 0.02]]))
    assert abs(float(cam.vfov) - 2 * np.arctan(480 / 1020)) < 1e-6
    s = cam.scale((0.5, 0.25))
    assert torch.allclose(s._data, torch.tensor([[320.0, 120.0, 250.0, 127.5, 160.0, 60.0, -0.1, 0.02]]))
    back = s.undo_scale_crop({"scales": torch.tensor([0.5, 0.25])})
    assert torch.allclose(back._data, cam._data)
    c = cam.crop((-40.0, 0.0))
    assert torch.allclose(c._data[0, [0, 4]], torch.tensor([600.0, 300.0]))
    p = G.Pinhole(torch.tensor([[640.0, 480.0, 500.0, 500.0, 320.0, 240.0]]))
    assert p._data.shape == (1, 8) and not hasattr(p, "dist")
    sr = G.SimpleRadial(p._data).update_dist(torch.tensor([[0.9]]))
    assert torch.allclose(sr.dist, torch.tensor([[0.7, 0.7]]))  # clamp + 2-wide dist block (quirk 3)
    f2 = p.update_focal(torch.tensor([[10.0]]), as_log=True)
    assert abs(float(f2.vfov) - np.deg2rad(5.0)) < 1e-5  # clamped to vfov >= 5 deg
    K = cam.K
    assert K.shape == (1, 3, 3) and float(K[0, 1, 2]) == 240.0
    assert G.camera_models["simple_divisional"].DIST_RANGE == (-3.0, 3.0)


def test_gravity_container():
    g = G.Gravity.from_rp(torch.tensor([0.2, -0.4]), torch.tensor([-0.3, 0.1]))
    assert torch.allclose(g.vec3d.norm(dim=-1), torch.ones(2))
    rp = g.rp.numpy()
    np.testing.assert_allclose(rp, oracle.gravity_rp(g.vec3d.numpy()), atol=1e-6)
    assert abs(rp[0, 0] - 0.2) < 5e-5 and abs(rp[0, 1] + 0.3) < 1e-6  # 1e-4 eps bias (quirk 10)
    J = g.J_rp()
    assert J.shape == (2, 3, 2)
    assert torch.allclose(torch.einsum("bij,bi->bj", J, g.vec3d), torch.zeros(2, 2), atol=1e-4)  # tangent


def test_tile_config_balances_jobs_for_any_width():
    """gclm_tile_config (host-only): every warp of the CTA gets the same number of (strip, band) jobs, the
    tiles cover the image, and gclm_reserve sizes its partial-sum scratch for at least that many tiles."""
    import ctypes
    from geocalib_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int32 * 5)()
    for model in range(4):
        for B in (1, 3, 64, 1024):
            for H in (1, 5, 37, 240, 480, 1024, 2000):
                for W in (4, 5, 36, 128, 320, 640, 768, 896, 1024, 1920, 4096):
                    for rows in (0, 1, 32):
                        assert lib.gclm_tile_config(B, H, W, model, 148, rows, out) == 0
                        nwarps, rsub, rpt, tiles, reserved = list(out)
                        ns = ((W + 3) // 4 * 4 + 127) // 128
                        assert 1 <= nwarps <= 10 and rsub >= 1 and rpt >= 1
                        assert (ns * rsub) % nwarps == 0, (model, B, H, W, rows, list(out))
                        assert rpt % rsub == 0 and tiles == (H + rpt - 1) // rpt
                        if rows == 0:
                            assert tiles <= reserved, (model, B, H, W, list(out))
    assert lib.gclm_tile_config(0, 480, 640, 0, 148, 0, out) < 0
    assert lib.gclm_tile_config(1, 480, 640, 7, 148, 0, out) < 0
