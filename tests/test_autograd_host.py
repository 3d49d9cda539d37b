"""Host-side pieces of the differentiable LM (geocalib_b200/autograd.py) on CPU: the local parameterisation of the
gravity direction against the reference's SphericalManifold (geocalib/misc.py:168-259), and the argument checks."""
import sys
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "baseline"))
import ref_import  # noqa: E402

from geocalib_b200 import autograd as AG  # noqa: E402


def unit(n, seed):
    g = torch.randn(n, 3, generator=torch.Generator().manual_seed(seed), dtype=torch.float64)
    g[0] = torch.tensor([0.0, 0.0, 1.0])    # the poles of the Householder construction
    g[1] = torch.tensor([0.0, 0.0, -1.0])
    g[2] = torch.tensor([1e-5, 0.0, -1.0])
    return torch.nn.functional.normalize(g, dim=-1)


def test_tangent_basis_is_the_derivative_of_plus():
    g = unit(16, 0)
    T = AG.tangent_basis(g)
    assert T.shape == (16, 3, 2)
    assert torch.allclose(torch.einsum("bij,bi->bj", T, g), torch.zeros(16, 2, dtype=torch.float64), atol=1e-12)  # tangent
    assert torch.allclose(T.transpose(1, 2) @ T, torch.eye(2, dtype=torch.float64).expand(16, 2, 2), atol=1e-12)  # orthonormal
    h = 1e-6
    for j in range(2):
        d = torch.zeros(16, 2, dtype=torch.float64)
        d[:, j] = h
        fd = (AG.sphere_plus(g, d) - AG.sphere_plus(g, -d)) / (2 * h)
        assert torch.allclose(fd, T[:, :, j], atol=1e-8), (fd - T[:, :, j]).abs().max()
    assert torch.allclose(AG.sphere_plus(g, torch.zeros(16, 2, dtype=torch.float64)), g, atol=1e-12)


@pytest.mark.skipif(not ref_import.available(), reason="baseline/_ref is not installed")
def test_parameterisation_equals_the_reference_manifold():
    ref_import.import_reference()
    from geocalib.misc import SphericalManifold

    g = unit(32, 1).float()
    d = 0.3 * torch.randn(32, 2, generator=torch.Generator().manual_seed(2))
    d[5] = 0.0
    assert torch.allclose(AG.sphere_plus(g, d), torch.nn.functional.normalize(SphericalManifold.plus(g, d), dim=-1), atol=2e-6)
    assert torch.allclose(AG.tangent_basis(g), SphericalManifold.J_plus(g), atol=2e-6)


def test_differentiable_calibrate_checks_its_arguments_before_touching_a_gpu():
    import geocalib_b200 as G
    from geocalib_b200 import _lib

    f = {"up_field": torch.zeros(1, 2, 8, 8), "latitude_field": torch.zeros(1, 1, 8, 8)}
    opt = G.LMOptimizer({"camera_model": "pinhole"}).eval()
    with pytest.raises(_lib.GclmError, match="up_confidence"):
        G.differentiable_calibrate(opt, f)
    full = dict(f, up_confidence=torch.ones(1, 8, 8), latitude_confidence=torch.ones(1, 8, 8))
    with pytest.raises(_lib.GclmError, match="prior_focal"):   # priors would silently be ignored otherwise
        G.differentiable_calibrate(opt, dict(full, prior_focal=torch.tensor([300.0])))
    opt.shared_intrinsics = True
    with pytest.raises(_lib.GclmError, match="per-image"):
        G.differentiable_calibrate(opt, full)
    opt.shared_intrinsics = False
    opt2 = G.LMOptimizer({"camera_model": "pinhole", "use_log_focal": False}).eval()
    with pytest.raises(_lib.GclmError, match="log focal"):
        G.differentiable_calibrate(opt2, full)


@pytest.mark.parametrize("K", [0, 1, 2])
def test_perturbed_parameters_are_the_retraction_of_the_lm_update(K):
    """theta (+) +-h e_j, batched: against the one-at-a-time formulation (sphere plus, f exp(delta), k + delta)."""
    B, h, n = 3, 1e-3, 3 + K
    g = unit(B, 4).float()
    cam = torch.tensor([[640.0, 480, 500 + 10 * b, 510 + 10 * b, 320, 240, 0.1 * (K > 0), -0.02 * (K > 1)] for b in range(B)])
    if K == 1:
        cam[:, 7] = cam[:, 6]
    cams, gravs = AG.perturbed_parameters(cam, g, K, h)
    assert cams.shape == (2 * n, B, 8) and gravs.shape == (2 * n, B, 3) and cams.is_contiguous() and gravs.is_contiguous()
    for j in range(n):
        for s, sign in enumerate((+1.0, -1.0)):
            d = torch.zeros(B, n)
            d[:, j] = sign * h
            gr = AG.sphere_plus(g, d[:, :2])
            cm = cam.clone()
            cm[:, 2:4] = cam[:, 2:4] * torch.exp(d[:, 2:3])
            if K >= 1:
                cm[:, 6] = cam[:, 6] + d[:, 3]
            if K == 2:
                cm[:, 7] = cam[:, 7] + d[:, 4]
            elif K == 1:
                cm[:, 7] = cm[:, 6]
            assert torch.allclose(cams[2 * j + s], cm, rtol=0, atol=0), (j, sign)
            assert torch.allclose(gravs[2 * j + s], gr, rtol=0, atol=1e-7), (j, sign)
    assert torch.equal(cams[:, :, [0, 1, 4, 5]], cam[:, [0, 1, 4, 5]].expand(2 * n, B, 4))  # size / principal point untouched
