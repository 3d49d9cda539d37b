"""`geocalib_b200.calibrate_batch` against a loop of the reference's own B == 1 `calibrate()`
(extractor.py:71-127).  The CPU test runs the UNMODIFIED reference end to end (its CNN with random weights, its
LMOptimizer) and checks that the batched call returns the same keys / types / shapes and -- up to the
reference's batch-global early stop -- the same values; the GPU test does the same with the engine installed,
where `stop_mode="per_image"` makes the batched result equal to the loop.
"""
import sys
from pathlib import Path

import numpy as np
import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "baseline"))
import ref_import  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="baseline/_ref is not installed")


def make_extractor(tmp_path, device):
    ref_import.import_reference()
    import geocalib
    from geocalib.geocalib import GeoCalib as Net

    torch.manual_seed(7)
    ckpt = tmp_path / "random.tar"
    torch.save({"model": Net().state_dict()}, ckpt)
    return geocalib.GeoCalib(weights=str(ckpt)).to(device).eval()


def blend(ex, camera_model):
    """Decoder output blended with a rendered field whose parameters depend on the image mean (so that every
    image has its own answer and a B == 1 call sees the same field as the batched call)."""
    from geocalib.camera import camera_models
    from geocalib.gravity import Gravity
    from geocalib.perspective_fields import get_perspective_field

    net = ex.model
    orig_forward = net.perspective_decoder.forward
    state = {}
    orig_ll = net.ll_enc.forward

    def ll_fwd(data):  # remember the images of this forward
        state["mean"] = data["image"].mean(dim=(1, 2, 3))
        return orig_ll(data)

    def fwd(data):
        out = orig_forward(data)
        B, _, h, w = out["up_field"].shape
        dev = out["up_field"].device
        m = state["mean"]
        f = 250.0 + 300.0 * (m - 0.45) * 10
        cam = torch.stack([torch.full_like(f, w), torch.full_like(f, h), f, f, torch.full_like(f, w / 2),
                           torch.full_like(f, h / 2), torch.zeros_like(f), torch.zeros_like(f)], -1)
        grav = Gravity.from_rp((m - 0.5) * 4, (0.5 - m) * 3).float()
        up, lat = get_perspective_field(camera_models[camera_model](cam), grav)
        out["up_field"] = torch.nn.functional.normalize(0.95 * up + 0.05 * out["up_field"], dim=1)
        out["latitude_field"] = 0.95 * lat + 0.05 * out["latitude_field"]
        return out

    net.perspective_decoder.forward = fwd
    net.ll_enc.forward = ll_fwd

    def restore():
        net.perspective_decoder.forward = orig_forward
        net.ll_enc.forward = orig_ll

    return restore


def images(device):
    g = torch.Generator().manual_seed(11)
    a = [torch.rand(3, 240, 320, generator=g) * s for s in (0.8, 0.9, 1.0)]
    b = [torch.rand(3, 256, 256, generator=g) * s for s in (0.85, 0.95)]
    return [x.to(device) for x in (a[0], b[0], a[1], b[1], a[2])]  # interleaved sizes: input order must survive


def check(batch, loop, tol_f, tol_g):
    n = len(loop)
    assert set(batch.keys()) == set(loop[0].keys())
    assert type(batch["camera"]).__name__ == type(loop[0]["camera"]).__name__ and batch["camera"]._data.shape == (n, 8)
    assert type(batch["gravity"]).__name__ == "Gravity" and batch["gravity"]._data.shape == (n, 3)
    assert batch["covariance"].shape == (n, 3, 3)
    for i, one in enumerate(loop):
        for k, v in one.items():
            if "uncertainty" in k:
                assert batch[k].shape == (n,), k
                np.testing.assert_allclose(float(batch[k][i]), float(v[0]), rtol=0.05, atol=1e-6, err_msg=k)
            elif "field" in k or "confidence" in k:
                got = batch[k][i]
                assert got.shape == v[0].shape, (k, got.shape, v.shape)  # back at the INPUT resolution
                assert torch.allclose(got, v[0], atol=1e-4), k
        f, fr = batch["camera"].f[i].cpu().numpy(), one["camera"].f[0].cpu().numpy()
        assert np.all(np.abs(f - fr) <= tol_f * np.abs(fr)), (i, f, fr)
        assert np.all(batch["camera"].size[i].cpu().numpy() == one["camera"].size[0].cpu().numpy())
        assert np.abs(batch["gravity"].vec3d[i].cpu().numpy() - one["gravity"].vec3d[0].cpu().numpy()).max() <= tol_g


def test_calibrate_batch_matches_a_loop_of_calibrate_reference_optimizer_cpu(tmp_path):
    import geocalib_b200 as G

    dev = torch.device("cpu")
    ex = make_extractor(tmp_path, dev)
    restore = blend(ex, "pinhole")
    try:
        imgs = images(dev)
        loop = [ex.calibrate(im) for im in imgs]
        batch = G.calibrate_batch(ex, imgs, max_batch=2)
        # the reference optimizer stops a batch when ALL its images converged (lm_optimizer.py:90-92): a few
        # extra iterations for some images -> agreement to 1e-3, not bit-wise
        check(batch, loop, tol_f=1e-3, tol_g=1e-3)
        assert isinstance(batch["up_field"], list) and len(batch["up_field"]) == 5  # two image sizes
        # priors are per image
        pf = torch.tensor([300.0, 310.0, 320.0, 330.0, 340.0])
        bp = G.calibrate_batch(ex, imgs, priors={"focal": pf})
        np.testing.assert_allclose(bp["camera"].f[:, 1].numpy(), pf.numpy(), rtol=1e-5)
        # same-size batch -> tensors; shared intrinsics is the reference's own batched path
        same = [imgs[0], imgs[2], imgs[4]]
        bs = G.calibrate_batch(ex, torch.stack(same), shared_intrinsics=True)
        rs = ex.calibrate(torch.stack(same), shared_intrinsics=True)
        assert torch.allclose(bs["camera"].f, rs["camera"].f) and bs["up_field"].shape == rs["up_field"].shape
    finally:
        restore()


@pytest.mark.gpu
def test_calibrate_batch_with_the_engine_equals_the_loop_gpu(tmp_path):
    import geocalib_b200 as G

    dev = torch.device("cuda", 0)
    ex = make_extractor(tmp_path, dev)
    restore = blend(ex, "simple_radial")
    try:
        G.install(ex)
        imgs = images(dev)
        loop = [ex.calibrate(im, camera_model="simple_radial") for im in imgs]
        eng = G.engine.get_engine(dev)
        l0 = eng.launch_count
        batch = G.calibrate_batch(ex, imgs, camera_model="simple_radial")
        assert eng.launch_count > l0
        assert ex.model.optimizer.conf.stop_mode == "batch"  # restored
        # per-image stop in the batched solve == what the B == 1 calls do; what remains is cuDNN picking other
        # convolution algorithms for other batch sizes (fields differ in the last bits)
        n = len(loop)
        for i, one in enumerate(loop):
            f, fr = batch["camera"].f[i].cpu().numpy(), one["camera"].f[0].cpu().numpy()
            assert np.all(np.abs(f - fr) <= 2e-4 * np.abs(fr)), (i, f, fr)
            k, kr = float(batch["camera"]._data[i, 6]), float(one["camera"]._data[0, 6])
            assert abs(k - kr) <= 2e-4
            assert np.abs(batch["gravity"].vec3d[i].cpu().numpy() - one["gravity"].vec3d[0].cpu().numpy()).max() <= 2e-4
        assert batch["covariance"].shape == (n, 4, 4)
    finally:
        restore()
