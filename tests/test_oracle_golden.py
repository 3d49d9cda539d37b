"""Pins the CPU oracle (oracle/gclm_oracle.c) against golden vectors produced by the reference
itself (oracle/make_golden.py, geocalib @ 3ee7ec34).  CPU only."""
import numpy as np
import pytest

from oracle import oracle

MODELS = ["pinhole", "simple_radial", "radial", "simple_divisional"]


def _load(golden_dir, name):
    return dict(np.load(golden_dir / name, allow_pickle=False))


def _rel(a, b):
    return np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)


@pytest.mark.parametrize("model", MODELS)
@pytest.mark.parametrize("tag,dtype,tol", [("f64", np.float64, 1e-10), ("f32", np.float32, 2e-4)])
def test_system_matches_reference(golden_dir, model, tag, dtype, tol):
    """G, H (both bases) and mean costs of setup_system / calculate_costs at fixed parameters."""
    g = _load(golden_dir, f"system_{model}.npz")
    B = g["cam"].shape[0]
    for b in range(B):
        for basis, rpf in (("opt", False), ("rpf", True)):
            G, H, cost = oracle.system(model, g["cam"][b], g["grav"][b], g["up_field"][b],
                                       g["latitude_field"][b], g["up_confidence"][b],
                                       g["latitude_confidence"][b], rpf_basis=rpf, dtype=dtype)
            Gr, Hr = g[f"G_{basis}_{tag}"][b], g[f"H_{basis}_{tag}"][b]
            assert _rel(G, Gr) < tol, (model, b, basis, G, Gr)
            assert _rel(H, Hr) < tol, (model, b, basis)
            assert abs(cost[0] - g[f"cost_up_{tag}"][b]) < tol * abs(g[f"cost_up_{tag}"][b])
            assert abs(cost[1] - g[f"cost_lat_{tag}"][b]) < tol * abs(g[f"cost_lat_{tag}"][b])


@pytest.mark.parametrize("model", MODELS)
def test_render_matches_reference(golden_dir, model):
    g = _load(golden_dir, f"system_{model}.npz")
    for b in range(g["cam_gt"].shape[0]):
        up, lat = oracle.render(model, g["cam_gt"][b], g["grav_gt"][b], 24, 32)
        assert np.abs(up - g["up_gt"][b]).max() < 2e-6
        assert np.abs(lat - g["lat_gt"][b]).max() < 2e-5


# ----------------------------------------------------------------------------------------------
# end-to-end LM: LMOptimizer(conf).eval().forward(data) golden outputs
# ----------------------------------------------------------------------------------------------
LM_CASES = [
    "default_pinhole", "default_simple_radial", "default_radial", "default_simple_divisional",
    "fixed10_pinhole", "fixlambda_pinhole", "euclid_pinhole", "linfocal_simple_radial",
    "scales_pinhole", "priorfocal_pinhole", "priorgravity_pinhole", "priorgravity_radial",
    "priordist_simple_radial", "latonly_pinhole", "noconf_simple_divisional",
    "shared_pinhole", "shared_radial", "shared_simple_divisional",
]


def run_oracle_on_golden(g, dtype=np.float32, **kw):
    model = str(g["model"])
    conf = {k: (v if k in ("lambda_", "atol", "rtol", "up_loss_fn_scale", "lat_loss_fn_scale") else int(v))
            for k, v in zip(g["conf_keys"].tolist(), g["conf_vals"].tolist())}
    conf.update(kw.pop("conf", {}))
    return oracle.lm(
        model, g.get("in_up_field"), g.get("in_latitude_field"), g.get("in_up_confidence"),
        g.get("in_latitude_confidence"), conf=conf, scales=g.get("scales"),
        prior_focal=g.get("prior_focal"), prior_gravity=g.get("prior_gravity"),
        prior_dist=g.get("prior_dist"), shared_intrinsics=bool(g["shared"]), dtype=dtype, **kw)


def assert_params_close(out, g, prefix="out_", tol=1e-4):
    """The north-star tolerance: 1e-4 relative on focal, gravity angles and distortion."""
    cam, cam_r = out["camera"], g[prefix + "camera"]
    K = oracle.NUM_DIST[str(g["model"])]
    assert np.all(np.abs(cam[:, 2:4] - cam_r[:, 2:4]) <= tol * np.abs(cam_r[:, 2:4])), (cam, cam_r)
    if K:
        assert np.all(np.abs(cam[:, 6:6 + K] - cam_r[:, 6:6 + K]) <= tol * np.maximum(1, np.abs(cam_r[:, 6:6 + K])))
    rp, rp_r = oracle.gravity_rp(out["gravity"]), oracle.gravity_rp(g[prefix + "gravity"])
    assert np.all(np.abs(rp - rp_r) <= tol * np.maximum(1, np.abs(rp_r))), (rp, rp_r)
    assert np.abs(out["gravity"] - g[prefix + "gravity"]).max() <= tol


@pytest.mark.parametrize("case", LM_CASES)
def test_lm_matches_reference(golden_dir, case):
    g = _load(golden_dir, f"lm_{case}.npz")
    out = run_oracle_on_golden(g)
    assert_params_close(out, g)
    np.testing.assert_array_equal(out["stop_at"], g["out_stop_at"].astype(np.int32))
    for k in ("initial_cost", "initial_up_cost", "initial_latitude_cost", "final_cost",
              "final_up_cost", "final_latitude_cost"):
        if "out_" + k in g:
            np.testing.assert_allclose(out[k], g["out_" + k], rtol=2e-4, atol=1e-9, err_msg=k)
    # covariance / uncertainties: inverse of an fp32 Hessian, compare loosely
    np.testing.assert_allclose(out["covariance"], g["out_covariance"], rtol=5e-3,
                               atol=5e-3 * np.abs(g["out_covariance"]).max())
    for k in ("roll_uncertainty", "pitch_uncertainty", "gravity_uncertainty", "focal_uncertainty",
              "vfov_uncertainty"):
        np.testing.assert_allclose(out[k], g["out_" + k], rtol=5e-3, atol=1e-7, err_msg=k)


@pytest.mark.parametrize("case", ["default_pinhole", "default_radial", "default_simple_divisional",
                                  "shared_pinhole", "priorgravity_radial"])
def test_lm_trajectory_matches_reference(golden_dir, case):
    """Parameters after exactly 1, 2, 3 LM iterations (catches errors convergence would hide)."""
    g = _load(golden_dir, f"lm_{case}.npz")
    for k in (1, 2, 3):
        out = run_oracle_on_golden(g, conf={"num_steps": k, "early_stop": False})
        assert_params_close(out, g, prefix=f"it{k}_", tol=2e-5)
        np.testing.assert_allclose(out["final_cost"], g[f"it{k}_final_cost"], rtol=2e-4)


# ----------------------------------------------------------------------------------------------
# BASELINE.json configurations at their real image sizes (reference outputs; seeded input recipe)
# ----------------------------------------------------------------------------------------------
BASE_CASES = ["config0_pinhole_320x240", "config1_pinhole_640x480", "config1_pinhole_640x480_default",
              "config2_simple_radial_640x480", "config3_simple_divisional_1024", "radial_640x480"]


def load_base_case(golden_dir, case):
    """Golden record + the regenerated inputs (checked against the stored digests)."""
    g = _load(golden_dir, f"base_{case}.npz")
    model = str(g["model"])
    f, cam_gt, _ = oracle.procedural_fields(model, int(g["B"]), int(g["h"]), int(g["w"]), int(g["seed"]))
    np.testing.assert_allclose(oracle.field_checksums(f), g["checksums"], rtol=1e-9,
                               err_msg="regenerated inputs differ from the ones the reference saw")
    np.testing.assert_array_equal(cam_gt, g["cam_gt"])
    conf = {k: (v if k in ("lambda_", "atol", "rtol") else int(v))
            for k, v in zip(g["conf_keys"].tolist(), g["conf_vals"].tolist())}
    return g, model, f, conf


@pytest.mark.parametrize("case", BASE_CASES)
def test_baseline_size_lm_matches_reference(golden_dir, case):
    g, model, f, conf = load_base_case(golden_dir, case)
    g = dict(g, shared=False)
    out = oracle.lm(model, f["up_field"], f["latitude_field"], f["up_confidence"], f["latitude_confidence"], conf=conf)
    assert_params_close(out, g)
    np.testing.assert_array_equal(out["stop_at"], g["out_stop_at"].astype(np.int32))
    np.testing.assert_allclose(out["final_cost"], g["out_final_cost"], rtol=2e-4)
    np.testing.assert_allclose(out["initial_cost"], g["out_initial_cost"], rtol=2e-4)
    for k in (1, 2, 3):
        o = oracle.lm(model, f["up_field"], f["latitude_field"], f["up_confidence"], f["latitude_confidence"],
                      conf={**conf, "num_steps": k, "early_stop": False})
        assert_params_close(o, g, prefix=f"it{k}_", tol=2e-5)


@pytest.mark.parametrize("case", ["config1_pinhole_640x480", "config2_simple_radial_640x480",
                                  "config3_simple_divisional_1024"])
def test_baseline_size_iteration0_system(golden_dir, case):
    """G, H and costs of iteration 0 (trivial initial estimate) in fp64 at full image size."""
    g, model, f, _ = load_base_case(golden_dir, case)
    for b in range(int(g["B"])):
        for basis, rpf in (("opt", False), ("rpf", True)):
            G, H, cost = oracle.system(model, g["init_camera"][b], g["init_gravity"][b], f["up_field"][b],
                                       f["latitude_field"][b], f["up_confidence"][b], f["latitude_confidence"][b],
                                       rpf_basis=rpf, dtype=np.float64)
            assert _rel(G, g[f"G0_{basis}_f64"][b]) < 1e-9
            assert _rel(H, g[f"H0_{basis}_f64"][b]) < 1e-9
        assert abs(cost[0] / g["cost0_up_f64"][b] - 1) < 1e-9 and abs(cost[1] / g["cost0_lat_f64"][b] - 1) < 1e-9
