"""Differentiable LM calibration.

The reference's training variant (`siclib/models/optimization/lm_optimizer.py:38-57, 577-625`) back-propagates
through 10 unrolled LM steps with autograd: every intermediate (B,N,2,n) Jacobian is kept alive (README: "smaller
batch size since lm takes more memory").  Here the forward pass is the engine's ordinary (inference) solve and
the backward pass differentiates the FIXED POINT instead of the iterations:

    G(theta*, x) = J^T W r = 0     =>     dL/dx = - d(lambda . G)/dx,    lambda = (dG/dtheta)^-T dL/dtheta

  * dL/dtheta: the incoming gradients on camera (fx, fy, k) and gravity, mapped to the LM's local parameters
    (spherical tangent, log focal) -- O(B) torch code;
  * dG/dtheta (B,n,n): central differences of `gclm_system` (the fused pass at perturbed parameters): 2 n passes --
    exact to second order, and unlike J^T W J it contains the second-derivative and Huber terms;
  * d(lambda . G)/dx: ONE streaming pass (`gclm_adjoint`) that writes the gradient of all four input planes.

Memory: the four gradient planes, nothing per iteration.  The gradient equals that of the unrolled reference once
the iterations have converged (tests/test_gpu_autograd.py compares with autograd through the UNMODIFIED reference
optimizer and with directional finite differences of the solve itself).  Per-image intrinsics, both fields present,
spherical manifold + log focal (the training configuration).
"""

from __future__ import annotations

from typing import Dict, Tuple

import torch

from . import _lib
from .engine import NUM_DIST, get_engine
from .lm_optimizer import LMOptimizer, get_trivial_estimation


def _householder(g: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """geocalib/misc.py:181-231 (the kernel's `householder`): v with v[2] = 1 and beta."""
    sigma = g[:, 0] ** 2 + g[:, 1] ** 2
    nrm = g.norm(dim=-1)
    sigma = torch.where(sigma < 1e-7, sigma + 1e-7, sigma)
    vpiv = torch.where(g[:, 2] < 0, g[:, 2] - nrm, -sigma / (g[:, 2] + nrm))
    beta = 2 * vpiv ** 2 / (sigma + vpiv ** 2)
    hv = torch.stack([g[:, 0] / vpiv, g[:, 1] / vpiv, torch.ones_like(vpiv)], -1)
    return hv, beta


def tangent_basis(g: torch.Tensor) -> torch.Tensor:
    """(B,3,2): d plus(g, delta) / d delta at delta = 0 (SphericalManifold.J_plus, misc.py:208-231)."""
    hv, beta = _householder(g)
    H = torch.eye(3, device=g.device, dtype=g.dtype)[None] - beta[:, None, None] * hv[:, :, None] * hv[:, None, :]
    return H[:, :, :2]


def sphere_plus(g: torch.Tensor, d: torch.Tensor) -> torch.Tensor:
    """SphericalManifold.plus (misc.py:233-259) followed by the normalisation of Gravity.__init__."""
    nd = d.norm(dim=-1)
    small = nd < 1e-7
    nd_ = torch.where(small, nd + 1e-7, nd)
    sinc = torch.where(small, torch.ones_like(nd), torch.sin(nd_) / nd_)
    ex = torch.cat([sinc[:, None] * d, torch.cos(nd)[:, None]], -1)
    hv, beta = _householder(g)
    out = g.norm(dim=-1, keepdim=True) * (ex - hv * (beta * (hv * ex).sum(-1))[:, None])
    return torch.nn.functional.normalize(out, dim=-1)


def perturbed_parameters(cam: torch.Tensor, grav: torch.Tensor, K: int, h: float) -> Tuple[torch.Tensor, torch.Tensor]:
    """theta (+) +-h e_j for all n = 3 + K local parameters [g0 g1 | log f | k..]: cameras (2n,B,8) and gravities
    (2n,B,3), ordered +e_0, -e_0, +e_1, ...  (the retraction of the LM update: sphere plus, f exp(delta), k + delta;
    for the 1-parameter models delta k1 also lands in the k2 slot)."""
    n, B = 3 + K, cam.shape[0]
    d = torch.zeros((2 * n, B, n), device=cam.device, dtype=cam.dtype)
    for j in range(n):
        d[2 * j, :, j] = h
        d[2 * j + 1, :, j] = -h
    gr = sphere_plus(grav.repeat(2 * n, 1), d[..., :2].reshape(2 * n * B, 2)).reshape(2 * n, B, 3)
    cm = cam.unsqueeze(0).repeat(2 * n, 1, 1)
    cm[..., 2:4] = cm[..., 2:4] * torch.exp(d[..., 2:3])
    if K >= 1:
        cm[..., 6] = cm[..., 6] + d[..., 3]
    if K == 2:
        cm[..., 7] = cm[..., 7] + d[..., 4]
    elif K == 1:
        cm[..., 7] = cm[..., 6]
    return cm.contiguous(), gr.contiguous()


class _LMFixedPoint(torch.autograd.Function):
    @staticmethod
    def forward(ctx, up, lat, uconf, lconf, optimizer: LMOptimizer, fd_step: float):
        data = {"up_field": up.detach(), "latitude_field": lat.detach(), "up_confidence": uconf.detach(),
                "latitude_confidence": lconf.detach()}
        with torch.no_grad():
            cam0, g0 = get_trivial_estimation(data, optimizer.camera_model)
            optimizer.setup_optimization_and_priors(data, shared_intrinsics=False)
            camera, gravity, infos = optimizer.optimize(data, cam0, g0)
        ctx.optimizer, ctx.fd_step = optimizer, fd_step
        ctx.save_for_backward(data["up_field"].contiguous(), data["latitude_field"].contiguous(),
                              data["up_confidence"].contiguous(), data["latitude_confidence"].contiguous(),
                              camera._data.detach(), gravity._data.detach())
        ctx.mark_non_differentiable(infos["stop_at"])
        return camera._data.clone(), gravity._data.clone(), infos["final_cost"].clone(), infos["stop_at"]

    @staticmethod
    def backward(ctx, g_cam, g_grav, g_cost, _g_stop):
        up, lat, uconf, lconf, cam, grav = ctx.saved_tensors
        opt: LMOptimizer = ctx.optimizer
        name = opt.camera_model.name()
        K, B = NUM_DIST[name], cam.shape[0]
        n = 3 + K
        eng = get_engine(up.device)
        options = opt._options()
        # ---- dL/dtheta in the local parameters [g0 g1 | log f | k..] ----
        T = tangent_basis(grav)                                         # (B,3,2)
        g_loc = torch.zeros((B, n), device=up.device, dtype=torch.float32)
        if g_grav is not None:
            g_loc[:, :2] = torch.einsum("bij,bi->bj", T, g_grav.float())
        if g_cam is not None:
            g_loc[:, 2] = cam[:, 2] * g_cam[:, 2] + cam[:, 3] * g_cam[:, 3]   # fx and fy both scale with exp(delta)
            if K == 1:
                g_loc[:, 3] = g_cam[:, 6] + g_cam[:, 7]  # the 2-wide dist block: delta k1 also lands in the k2 slot (quirk 3)
            elif K == 2:
                g_loc[:, 3:5] = g_cam[:, 6:8]

        # ---- dG/dtheta by central differences of the fused pass ----
        # all 2n perturbed parameter sets theta (+) +-h e_j in ONE batch of O(n B) torch work, then 2n system passes
        h = ctx.fd_step
        cams, gravs = perturbed_parameters(cam, grav, K, h)              # (2n, B, 8), (2n, B, 3)
        Gs = [eng.system(name, up, lat, uconf, lconf, cams[i], gravs[i], options=options, rpf_basis=False)[0]
              for i in range(2 * n)]
        Gs = torch.stack(Gs, 0).double()                                  # (2n, B, n): +h e_0, -h e_0, +h e_1, ...
        A = ((Gs[0::2] - Gs[1::2]) / (2 * h)).permute(1, 2, 0).contiguous()   # A[:, i, j] = dG_i / dtheta_j
        lam = torch.linalg.solve(A.transpose(1, 2), g_loc.double()).float()   # (dG/dtheta)^-T dL/dtheta
        g_up, g_lat, g_uc, g_lc = eng.adjoint(name, up, lat, uconf, lconf, cam, grav, lam, options=options)
        # G here is J^T W r with r = obs - pred (the LM ascent direction): the fixed point moves by -A^-1 dG/dx
        return -g_up, -g_lat, -g_uc, -g_lc, None, None


def differentiable_calibrate(optimizer: LMOptimizer, data: Dict[str, torch.Tensor], fd_step: float = 1e-3):
    """LM calibration whose outputs carry gradients w.r.t. the four field tensors.

    Returns (camera_data (B,8), gravity (B,3), final_cost (B,), stop_at (B,)); `camera_data[:, 2:4]` (focal),
    `[:, 6:8]` (distortion) and `gravity` are differentiable.  `final_cost` and `stop_at` are
    reported values: a gradient arriving on them is NOT propagated (use the parameters in the loss).  Run enough iterations to converge (the gradient is
    that of the fixed point): `num_steps` >= 20 with early stop is the usual choice."""
    for k in ("up_field", "latitude_field", "up_confidence", "latitude_confidence"):
        if k not in data:
            raise _lib.GclmError(f"differentiable_calibrate needs {k!r} (both fields with confidences)")
    extra = [k for k in ("prior_focal", "prior_gravity", "prior_dist", "scales", "packed_fields") if k in data]
    if extra:  # they change which parameters are free / the initial estimate: not wired into the backward yet
        raise _lib.GclmError(f"differentiable_calibrate: {extra} are not supported (all of gravity, focal, distortion are estimated)")
    if optimizer.shared_intrinsics:
        raise _lib.GclmError("differentiable_calibrate: per-image intrinsics only")
    c = optimizer.conf
    if not (c.use_spherical_manifold and c.use_log_focal):
        raise _lib.GclmError("differentiable_calibrate: spherical manifold + log focal (the training configuration)")
    return _LMFixedPoint.apply(data["up_field"], data["latitude_field"], data["up_confidence"],
                               data["latitude_confidence"], optimizer, float(fd_step))
