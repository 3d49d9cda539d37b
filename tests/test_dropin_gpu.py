"""The real drop-in, executed on a GPU: the UNMODIFIED reference package (installed in the git-ignored
baseline/_ref by baseline/install_ref.sh, so it travels to the GPU box) runs its own
`GeoCalib.calibrate()` (extractor.py:71-127) twice on the same images -- once with its own
`LMOptimizer` (lm_optimizer.py, PyTorch eager on CUDA) and once after `geocalib_b200.install()` swapped
in the sm_100a engine (geocalib.py:106,119) -- and the two result dicts are compared at 1e-4.

There is no network, hence no released checkpoint: the CNN has random weights.  So that the LM problem
is well-posed, the decoder's output fields are blended with a rendered perspective field (90 % rendered
+ 10 % CNN, re-normalised); confidences are the CNN's own.  Everything else is the reference's code
path: ImagePreprocessor -> MSCAN backbone -> decoders -> optimizer -> _post_process.
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "baseline"))
import ref_import  # noqa: E402

needs_ref = pytest.mark.skipif(not ref_import.available(),
                               reason="baseline/_ref is not installed (run baseline/install_ref.sh in the build container)")


@pytest.fixture(scope="module")
def extractor(tmp_path_factory):
    assert torch.cuda.is_available()
    ref_import.import_reference()
    import geocalib
    from geocalib.geocalib import GeoCalib as Net

    torch.manual_seed(7)
    net = Net()
    ckpt = tmp_path_factory.mktemp("w") / "random.tar"
    torch.save({"model": net.state_dict()}, ckpt)
    ex = geocalib.GeoCalib(weights=str(ckpt)).cuda().eval()  # extractor.py:36-43 (custom weights path)
    return ex


def blend_decoder(ex, camera_model, gt):
    """Wrap the perspective decoder: CNN output blended with the perspective field of `gt` cameras."""
    from geocalib.camera import camera_models
    from geocalib.gravity import Gravity
    from geocalib.perspective_fields import get_perspective_field

    dec = ex.model.perspective_decoder
    orig = dec.forward
    cache = {}  # cuDNN picks its convolution algorithms per call: replay the SAME decoder output to both optimizers

    def fwd(data):
        key = tuple(data["features"]["ll"].shape)
        if key in cache:
            return {k: v.clone() for k, v in cache[key].items()}
        out = orig(data)
        B, _, h, w = out["up_field"].shape
        dev = out["up_field"].device
        cam = torch.tensor([[w, h, gt["f"][b], gt["f"][b], w / 2, h / 2, gt["k1"][b], 0.0] for b in range(B)],
                           dtype=torch.float32, device=dev)
        grav = Gravity.from_rp(torch.tensor(gt["roll"][:B], device=dev), torch.tensor(gt["pitch"][:B], device=dev)).float()
        up, lat = get_perspective_field(camera_models[camera_model](cam), grav)
        u = 0.9 * up + 0.1 * out["up_field"]
        out["up_field"] = torch.nn.functional.normalize(u, dim=1)
        out["latitude_field"] = 0.9 * lat + 0.1 * out["latitude_field"]
        cache[key] = {k: v.clone() for k, v in out.items()}
        return out

    dec.forward = fwd
    return lambda: setattr(dec, "forward", orig)


def _np(x):
    return x.detach().cpu().numpy().astype(np.float64)


def compare(res, ref, K, tol=1e-4):
    assert set(res.keys()) == set(ref.keys())
    for k in ref:
        a, b = res[k], ref[k]
        if hasattr(b, "_data"):
            assert type(a).__name__ == type(b).__name__ and a._data.shape == b._data.shape, k
        else:
            assert a.shape == b.shape and a.dtype == b.dtype and a.device == b.device, k
    cf, cr = _np(res["camera"].f), _np(ref["camera"].f)
    assert np.all(np.abs(cf - cr) <= tol * np.abs(cr)), (cf, cr)
    if K:
        d, dr = _np(res["camera"]._data[:, 6:6 + K]), _np(ref["camera"]._data[:, 6:6 + K])
        assert np.all(np.abs(d - dr) <= tol * np.maximum(1, np.abs(dr))), (d, dr)
    assert np.abs(_np(res["gravity"].vec3d) - _np(ref["gravity"].vec3d)).max() <= tol
    rp, rpr = _np(res["gravity"].rp), _np(ref["gravity"].rp)  # the SAME reference property on both
    assert np.all(np.abs(rp - rpr) <= tol * np.maximum(1, np.abs(rpr)))
    for k in ref:
        if "field" in k or "confidence" in k:
            assert torch.equal(res[k], ref[k]), k  # CNN outputs: untouched by the swap
        if "uncertainty" in k:
            np.testing.assert_allclose(_np(res[k]), _np(ref[k]), rtol=5e-3, atol=1e-7, err_msg=k)
    np.testing.assert_allclose(_np(res["covariance"]), _np(ref["covariance"]), rtol=5e-3,
                               atol=5e-3 * np.abs(_np(ref["covariance"])).max())


@needs_ref
@pytest.mark.parametrize("camera_model,K", [("pinhole", 0), ("simple_radial", 1), ("simple_divisional", 1)])
def test_calibrate_single_image_through_the_reference_extractor(extractor, camera_model, K):
    import geocalib_b200 as G

    ex = extractor
    gt = {"f": [330.0], "k1": [-0.08 if K else 0.0], "roll": [0.15], "pitch": [-0.2]}
    restore = blend_decoder(ex, camera_model, gt)
    ref_opt = ex.model.optimizer
    try:
        img = torch.rand(3, 480, 640, generator=torch.Generator().manual_seed(3)).cuda()
        ref = ex.calibrate(img, camera_model=camera_model)
        new = G.install(ex)
        assert isinstance(ex.model.optimizer, G.LMOptimizer)
        eng = G.engine.get_engine(img.device)
        l0 = eng.launch_count
        res = ex.calibrate(img, camera_model=camera_model)
        torch.cuda.synchronize()
        assert eng.launch_count > l0 and eng.persistent_count > 0, "the engine did not run"  # B = 1: ONE cooperative launch
        compare(res, ref, K)
        assert abs(float(res["camera"].f[0, 1]) / (330.0 * 480 / 320) - 1) < 0.25  # undo_scale_crop applied
    finally:
        restore()
        ex.model.optimizer = ref_opt


@needs_ref
def test_calibrate_shared_intrinsics_and_priors_through_the_reference_extractor(extractor):
    import geocalib_b200 as G

    ex = extractor
    gt = {"f": [300.0] * 3, "k1": [0.0] * 3, "roll": [0.1, -0.2, 0.05], "pitch": [-0.1, 0.15, 0.3]}
    restore = blend_decoder(ex, "pinhole", gt)
    ref_opt = ex.model.optimizer
    try:
        imgs = torch.rand(3, 3, 384, 512, generator=torch.Generator().manual_seed(5)).cuda()
        ref = ex.calibrate(imgs, shared_intrinsics=True)
        ref_p = ex.calibrate(imgs[0], priors={"focal": torch.tensor(480.0).cuda()})
        G.install(ex)
        res = ex.calibrate(imgs, shared_intrinsics=True)
        res_p = ex.calibrate(imgs[0], priors={"focal": torch.tensor(480.0).cuda()})
        compare(res, ref, 0)
        assert float(_np(res["camera"].f[:, 1]).std()) < 1e-3  # one focal for all three images
        compare(res_p, ref_p, 0)
    finally:
        restore()
        ex.model.optimizer = ref_opt


@needs_ref
def test_forward_of_the_network_with_the_engine_installed(extractor):
    """geocalib.geocalib.GeoCalib.forward (geocalib.py:106-119) on an already preprocessed batch: the
    engine is called with B > 1 (what `calibrate` never does for per-image intrinsics)."""
    import geocalib_b200 as G

    ex = extractor
    B = 4
    gt = {"f": [280.0, 320.0, 400.0, 250.0], "k1": [0.0] * B, "roll": [0.1, -0.2, 0.05, 0.3], "pitch": [-0.1, 0.15, 0.3, 0.0]}
    restore = blend_decoder(ex, "pinhole", gt)
    ref_opt = ex.model.optimizer
    try:
        data = {"image": torch.rand(B, 3, 320, 416, generator=torch.Generator().manual_seed(9)).cuda()}
        ex.model.optimizer.set_camera_model("pinhole")
        ex.model.optimizer.shared_intrinsics = False
        with torch.no_grad():
            ref = ex.model(dict(data))
            G.install(ex)
            res = ex.model(dict(data))
        for k in ("camera", "gravity"):
            assert type(res[k]).__name__ == type(ref[k]).__name__
        cf, cr = _np(res["camera"].f), _np(ref["camera"].f)
        assert np.all(np.abs(cf - cr) <= 1e-4 * np.abs(cr)), (cf, cr)
        assert np.abs(_np(res["gravity"].vec3d) - _np(ref["gravity"].vec3d)).max() <= 1e-4
        assert np.all(np.abs(cf[:, 1] / np.array(gt["f"]) - 1) < 0.25)
    finally:
        restore()
        ex.model.optimizer = ref_opt
