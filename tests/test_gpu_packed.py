"""Pre-transformed field layout (include/gclm.h: gclm_pack_fields / gclm_pack_decoder_outputs /
gclm_solve_packed): [up_x, up_y, sin_lat, w_up, w_lat] with fp32 or fp16 confidences.

  * fp32 confidences: bit-identical to the plain solve (same numbers, no sin pass);
  * fp16 confidences (16 B/pixel/pass): MEASURED deviation of the estimates from the fp32 solve and from the
    oracle at the BASELINE image sizes -- required <= 1e-4 like everything else;
  * the fused decoder tail: from PRE-activation head outputs == torch's F.normalize / asin(clamp(tanh)) /
    sigmoid followed by the plain solve (geocalib.py:44-89).
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import oracle  # noqa: E402

from test_gpu_parity import assert_params_close, dev, render_on_device, synth_params  # noqa: E402,F401


def _np(out):
    return out["camera"]._data.cpu().numpy(), out["gravity"]._data.cpu().numpy()


@pytest.mark.parametrize("model,B,h,w", [("pinhole", 16, 480, 640), ("simple_radial", 8, 480, 640),
                                         ("simple_divisional", 2, 1024, 1024), ("radial", 4, 96, 128)])
def test_packed_layout_fp32_is_bit_identical_and_fp16_within_tolerance(dev, model, B, h, w):
    import geocalib_b200 as G
    from geocalib_b200.engine import get_engine

    cam_gt, grav_gt = oracle.gt_params(model, B, h, w, 2234)
    eng = get_engine(dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    f = eng.render_fields(model, t(cam_gt), t(grav_gt), h, w, sigma=0.02, conf_lo=0.2, seed=77)
    conf = {"camera_model": model, "num_steps": 20, "early_stop": False, "schedule": "per_pass"}
    with torch.no_grad():
        plain = G.LMOptimizer(conf).eval()(f)
        p32 = eng.pack_fields(f["up_field"], f["latitude_field"], f["up_confidence"], f["latitude_confidence"],
                              conf_dtype=torch.float32)
        out32 = G.LMOptimizer(conf).eval()({"packed_fields": p32})
        out16 = G.LMOptimizer({**conf, "confidence_dtype": "float16"}).eval()(f)
    torch.cuda.synchronize()
    assert p32["up_conf"].dtype == torch.float32
    for k in ("camera", "gravity"):
        assert torch.equal(plain[k]._data, out32[k]._data), k
    assert torch.equal(plain["final_cost"], out32["final_cost"]) and torch.equal(plain["covariance"], out32["covariance"])
    assert set(plain.keys()) == set(out32.keys())
    # fp16 confidences: weights perturbed by <= 2^-11 relative, zero-mean over 3e5 .. 1e6 pixels
    cam, grav = _np(out16)
    camp, gravp = _np(plain)
    assert_params_close(cam, grav, camp, gravp, model, tol=1e-4)
    dev_f = np.abs(cam[:, 3] / camp[:, 3] - 1).max()
    dev_g = np.abs(grav - gravp).max()
    print(f"\n[packed fp16 {model} {h}x{w}] max |df/f| = {dev_f:.2e}, max |dg| = {dev_g:.2e}, "
          f"max |dk| = {np.abs(cam[:, 6:] - camp[:, 6:]).max():.2e}")
    assert dev_f < 2e-5 and dev_g < 2e-5  # measured ~1e-6: two orders inside the tolerance
    if h * w <= 480 * 640:
        fn = {k: v.cpu().numpy() for k, v in f.items()}
        ref = oracle.lm(model, fn["up_field"], fn["latitude_field"], fn["up_confidence"], fn["latitude_confidence"],
                        conf={"num_steps": 20, "early_stop": False})
        assert_params_close(cam, grav, ref["camera"], ref["gravity"], model, tol=1e-4)


@pytest.mark.parametrize("conf_dtype", [torch.float32, torch.float16])
def test_fused_decoder_tail_matches_torch_activations(dev, conf_dtype):
    """Heads' pre-activation outputs -> gclm_pack_decoder_outputs -> gclm_solve_packed  ==  the reference's
    activations in torch (geocalib.py:57,73-75,86-88) -> plain solve."""
    import torch.nn.functional as F

    import geocalib_b200 as G
    from geocalib_b200.engine import get_engine

    model, B, h, w = "pinhole", 6, 96, 128
    cam, grav = synth_params(model, B, h, w, 41)
    f = render_on_device(model, cam, grav, h, w, dev, sigma=0.02, seed=9)
    # invert the activations to get plausible head outputs
    up_logits = (f["up_field"] * torch.empty(B, 1, h, w, device=dev).uniform_(0.5, 3.0)).contiguous()
    lat_logits = torch.atanh(torch.sin(f["latitude_field"]).clamp(-0.999, 0.999)).contiguous()
    ucl = torch.logit(f["up_confidence"].clamp(1e-3, 1 - 1e-3)).contiguous()
    lcl = torch.logit(f["latitude_confidence"].clamp(1e-3, 1 - 1e-3)).contiguous()
    eps = 1e-5
    activated = {"up_field": F.normalize(up_logits, dim=1),
                 "latitude_field": torch.asin(torch.clamp(torch.tanh(lat_logits), -1 + eps, 1 - eps)),
                 "up_confidence": torch.sigmoid(ucl), "latitude_confidence": torch.sigmoid(lcl)}
    eng = get_engine(dev)
    packed = eng.pack_decoder_outputs(up_logits, lat_logits, ucl, lcl, conf_dtype=conf_dtype)
    assert packed["up_conf"].dtype == conf_dtype
    assert torch.allclose(packed["up"], activated["up_field"], atol=1e-6)
    assert torch.allclose(packed["sin_lat"], torch.sin(activated["latitude_field"]), atol=2e-6)
    assert torch.allclose(packed["up_conf"].float(), activated["up_confidence"], atol=1e-3 if conf_dtype == torch.float16 else 1e-6)
    with torch.no_grad():
        ref = G.LMOptimizer({"camera_model": model}).eval()(activated)
        out = G.LMOptimizer({"camera_model": model}).eval()({"packed_fields": packed})
    torch.cuda.synchronize()
    assert_params_close(*_np(out), *_np(ref), model, tol=1e-4 if conf_dtype == torch.float16 else 2e-5)


def test_packed_rejects_what_it_cannot_stream(dev):
    import geocalib_b200 as G
    from geocalib_b200 import _lib
    from geocalib_b200.engine import get_engine

    cam, grav = synth_params("pinhole", 2, 20, 50, 1)  # width not a multiple of 4
    f = render_on_device("pinhole", cam, grav, 20, 50, dev, sigma=0.0, seed=1)
    eng = get_engine(dev)
    p = {"up": f["up_field"], "sin_lat": torch.sin(f["latitude_field"]), "up_conf": f["up_confidence"].half(),
         "lat_conf": f["latitude_confidence"].half()}
    with pytest.raises(_lib.GclmError):
        G.LMOptimizer({"camera_model": "pinhole"}).eval()({"packed_fields": p})
    # the drop-in silently keeps the fp32 path for such shapes (confidence_dtype is a hint, not a contract)
    with torch.no_grad():
        out = G.LMOptimizer({"camera_model": "pinhole", "confidence_dtype": "float16"}).eval()(f)
    assert torch.isfinite(out["camera"]._data).all()
