"""CUDA-graph front end for the reference CNN (geocalib_b200/frontend.py, SURVEY 8f.4 first step).

CPU part (build container / any box with baseline/_ref): the wiring -- `install_frontend(model, graph=False)` runs the
same statements as `GeoCalib.forward` and, with the same seed, the same random NMF bases, so `calibrate()` is bit-equal
to the untouched reference.  GPU part: the captured graph against the eager modules, fresh random bases on every
replay, and `calibrate()` through graph + engine.
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


def fixed_bases(self, B, S, D, R, device="cpu"):
    """Deterministic stand-in for the random NMF bases (same for eager and graph runs)."""
    i = torch.arange(B * S * D * R, device=device, dtype=torch.float32).reshape(B * S, D, R)
    return torch.nn.functional.normalize(0.5 + 0.5 * torch.sin(0.37 * i), dim=1)


def test_frontend_wiring_equals_the_reference_forward_on_cpu(tmp_path):
    import geocalib_b200 as G
    from geocalib_b200 import frontend as FE

    ex = make_extractor(tmp_path, "cpu")
    img = torch.rand(3, 96, 128, generator=torch.Generator().manual_seed(1))
    assert len(FE.nmf_modules(ex.model)) == 2  # one Hamburger per decoder head
    torch.manual_seed(11)
    ref = ex.calibrate(img)
    front = G.install_frontend(ex, graph=False)
    assert "forward" in ex.model.__dict__
    torch.manual_seed(11)
    res = ex.calibrate(img)  # same CPU generator, same draws: bit-equal
    for k in ref:
        a, b = (res[k]._data, ref[k]._data) if hasattr(ref[k], "_data") else (res[k], ref[k])
        assert torch.equal(a, b), k
    assert front.replays == 0 and front.captures == 0
    front.capturable_eager = True      # the capture-safe path (channel flip through the proxy): same bits
    torch.manual_seed(11)
    res = ex.calibrate(img)
    for k in ref:
        a, b = (res[k]._data, ref[k]._data) if hasattr(ref[k], "_data") else (res[k], ref[k])
        assert torch.equal(a, b), k
    G.uninstall_frontend(ex)
    assert "forward" not in ex.model.__dict__ and all("_build_bases" not in m.__dict__ for m in FE.nmf_modules(ex.model))
    torch.manual_seed(11)
    again = ex.calibrate(img)
    assert torch.equal(again["camera"]._data, ref["camera"]._data)
    with pytest.raises(RuntimeError, match="CUDA tensors only"):
        FE.GraphedFrontEnd(ex.model)(img[None])


def test_graph_cache_is_bounded_and_least_recently_used():
    """Pure host logic (no reference, no GPU): one entry per shape, capped, oldest dropped first."""
    from geocalib_b200 import frontend as FE

    front = FE.GraphedFrontEnd(torch.nn.Identity(), max_graphs=2)
    made = []
    front._capture = lambda image: made.append(tuple(image.shape)) or ("graph", image, {"shape": tuple(image.shape)})
    a, b, c = torch.zeros(1, 3, 32, 32), torch.zeros(1, 3, 64, 32), torch.zeros(2, 3, 32, 32)
    key = lambda t: (None, tuple(t.shape), t.dtype)
    for t in (a, b, a, c, a, b):            # a, b captured; a re-used; c evicts b; a re-used; b captured again, evicts c
        assert front._lookup(key(t), t)[2]["shape"] == tuple(t.shape)
    assert made == [tuple(a.shape), tuple(b.shape), tuple(c.shape), tuple(b.shape)]
    assert list(front._graphs) == [key(a), key(b)]


@pytest.mark.gpu
def test_graph_replay_equals_the_eager_cnn_and_redraws_the_bases(tmp_path):
    from geocalib_b200 import frontend as FE

    ex = make_extractor(tmp_path, "cuda")
    net = ex.model
    img = torch.rand(1, 3, 320, 448, generator=torch.Generator().manual_seed(2)).cuda()
    undo = FE.use_device_bases(net, fixed_bases)
    try:
        with torch.no_grad():
            eager = {k: v.clone() for k, v in FE.cnn_forward(net, img).items()}   # the reference's statements, eager
        front = FE.GraphedFrontEnd(net)
        out1 = {k: v.clone() for k, v in front(img).items()}
        out2 = {k: v.clone() for k, v in front(img).items()}
        assert front.captures == 1 and front.replays == 2
        for k in eager:
            assert torch.equal(out1[k], out2[k]), k                      # a replay is deterministic
            d = float((out1[k] - eager[k]).abs().max())
            print(f"\n[frontend] {k}: max |graph - eager| = {d:.3e}")
            assert d < 1e-4, (k, d)                                       # same kernels up to cuDNN's algorithm choice
        other = front(torch.rand(1, 3, 320, 448, generator=torch.Generator().manual_seed(3)).cuda())
        assert float((other["latitude_field"] - out1["latitude_field"]).abs().max()) > 1e-4   # static input is refreshed
        front(torch.rand(1, 3, 256, 320).cuda())
        assert front.captures == 2                                        # one graph per shape
    finally:
        undo()
    # the default (random) bases are drawn again on every replay: the CUDA generator advances under the graph
    undo = FE.use_device_bases(net)
    try:
        front = FE.GraphedFrontEnd(net)
        a = front(img)["up_confidence"].clone()
        b = front(img)["up_confidence"].clone()
        assert not torch.equal(a, b)
    finally:
        undo()


@pytest.mark.gpu
def test_calibrate_through_graph_and_engine(tmp_path):
    import geocalib_b200 as G
    from geocalib_b200 import frontend as FE
    from test_dropin_gpu import blend_decoder, compare

    ex = make_extractor(tmp_path, "cuda")
    gt = {"f": [330.0], "k1": [0.0], "roll": [0.15], "pitch": [-0.2]}
    img = torch.rand(3, 480, 640, generator=torch.Generator().manual_seed(3)).cuda()
    restore = blend_decoder(ex, "pinhole", gt)   # well-posed fields from a random-weight CNN; caches the decoder output
    try:
        ref = ex.calibrate(img)                   # reference network, reference optimizer, eager
        G.install(ex)
        front = G.install_frontend(ex)
        res = ex.calibrate(img)
        res2 = ex.calibrate(img)
        torch.cuda.synchronize()
        assert front.captures == 1 and front.replays == 2
        compare(res, ref, 0)
        compare(res2, ref, 0)
    finally:
        G.uninstall_frontend(ex)
        restore()
