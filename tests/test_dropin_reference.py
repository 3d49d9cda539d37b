"""Drop-in seam against the REAL reference classes (CPU, build container only).

The reference lives at /root/reference in the build container and nowhere else, so this module is
skipped when it is absent (e.g. on the GPU box).  It checks the plumbing of the swap, not numerics:
`install()` replaces `GeoCalib.optimizer` (geocalib/geocalib.py:106) without touching weights, the
replacement answers everything `extractor.calibrate` pokes (extractor.py:112-113), results would be
wrapped in the reference's own Camera/Gravity classes, and on CPU tensors the engine fails LOUDLY
instead of falling back.
"""
import sys
import types
from pathlib import Path

import pytest
import torch

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "geocalib" / "lm_optimizer.py").exists(),
                                reason="reference checkout not available on this machine")


@pytest.fixture(scope="module")
def ref_modules():
    sys.modules.setdefault("kornia", types.ModuleType("kornia"))  # only used by ImagePreprocessor.resize
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))
    import geocalib.camera as rcam
    import geocalib.geocalib as rnet
    import geocalib.gravity as rgrav
    import geocalib.lm_optimizer as rlm

    return types.SimpleNamespace(cam=rcam, net=rnet, grav=rgrav, lm=rlm)


def test_install_swaps_the_optimizer_of_the_reference_network(ref_modules):
    import geocalib_b200 as G

    net = ref_modules.net.GeoCalib(num_steps=12, lambda_=0.05).eval()
    keys_before = set(net.state_dict().keys())
    ref_opt = net.optimizer
    assert isinstance(ref_opt, ref_modules.lm.LMOptimizer)
    new = G.install(net)
    assert net.optimizer is new and isinstance(new, G.LMOptimizer)
    assert set(net.state_dict().keys()) == keys_before  # the LM has no weights: checkpoints unaffected
    assert (new.num_steps, new.conf.lambda_, new.training) == (12, 0.05, False)
    for k, v in vars(ref_opt.conf).items():  # all 13 reference keys carried over
        assert getattr(new.conf, k) == v, k
    # what extractor.calibrate does before every forward (extractor.py:112-113)
    net.optimizer.set_camera_model("simple_divisional")
    net.optimizer.shared_intrinsics = True
    assert net.optimizer.camera_model.name() == "simple_divisional" and net.optimizer.camera_has_distortion
    # same bookkeeping as the reference for every prior combination
    for data in ({}, {"prior_focal": 1}, {"prior_gravity": 1}, {"prior_dist": 1}, {"prior_gravity": 1, "prior_focal": 1}):
        ref_opt.set_camera_model("simple_divisional")
        ref_opt.setup_optimization_and_priors(dict(data), shared_intrinsics=False)
        new.setup_optimization_and_priors(dict(data), shared_intrinsics=False)
        assert (new.estimate_gravity, new.estimate_focal, bool(new.estimate_dist)) == \
               (ref_opt.estimate_gravity, ref_opt.estimate_focal, bool(ref_opt.estimate_dist)), data
        assert new.n_intrinsic_params == ref_opt.n_intrinsic_params
        assert new.gravity_delta_dims == ref_opt.gravity_delta_dims and new.focal_delta_dims == ref_opt.focal_delta_dims


def test_results_are_wrapped_in_the_reference_classes(ref_modules):
    from geocalib_b200.camera import wrap_camera
    from geocalib_b200.gravity import wrap_gravity

    cam = wrap_camera("radial", torch.tensor([[640.0, 480.0, 500.0, 500.0, 320.0, 240.0, -0.1, 0.01]]))
    grav = wrap_gravity(torch.tensor([[0.1, -0.9, 0.2]]))
    assert isinstance(cam, ref_modules.cam.Radial) and isinstance(grav, ref_modules.grav.Gravity)
    # ... and our mirrors agree with them on everything the callers read
    import geocalib_b200 as G

    mine = G.Radial(cam._data.clone())
    for attr in ("size", "f", "c", "dist", "vfov", "hfov", "K", "k1", "k2"):
        assert torch.allclose(getattr(mine, attr), getattr(cam, attr)), attr
    data = {"scales": torch.tensor([0.5, 0.4]), "crop_pad": torch.tensor([-8.0, 0.0])}
    assert torch.allclose(mine.undo_scale_crop(data)._data, cam.undo_scale_crop(data)._data)
    d = torch.tensor([[0.05]])
    assert torch.allclose(mine.update_focal(d, as_log=True)._data, cam.update_focal(d, as_log=True)._data)
    assert torch.allclose(mine.update_dist(torch.tensor([[0.9, -0.9]]))._data, cam.update_dist(torch.tensor([[0.9, -0.9]]))._data)
    g_mine = G.Gravity(grav._data.clone())
    assert torch.allclose(g_mine.rp, grav.rp, atol=1e-6) and torch.allclose(g_mine.J_rp(), grav.J_rp(), atol=1e-6)


def test_forward_on_cpu_fails_loudly_through_the_reference_network(ref_modules):
    """The reference network runs its CNN unchanged and hands the predicted fields to the swapped
    optimizer (geocalib.py:119); without a GPU the engine must raise, never compute on the CPU."""
    import geocalib_b200 as G
    from geocalib_b200._lib import GclmError

    net = ref_modules.net.GeoCalib().eval()
    G.install(net)
    img = torch.rand(1, 3, 64, 96)
    with torch.no_grad(), pytest.raises(GclmError, match="CUDA"):
        net({"image": img})
