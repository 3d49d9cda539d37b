"""Drop-in replacement for `geocalib.lm_optimizer.LMOptimizer` backed by the sm_100a engine.

Boundary being replaced (reference @ 3ee7ec34):
  * construction / attributes   geocalib/lm_optimizer.py:141-246
  * `forward(data) -> dict`     geocalib/lm_optimizer.py:646-664 (and :551-644, :463-516)
  * owner / caller              geocalib/geocalib.py:106,119; geocalib/extractor.py:112-113

Same class name, same 13 configuration keys, same mutable attributes (`shared_intrinsics`,
`set_camera_model`), same input dict, same output dict keys and shapes.  What differs is where the
work happens: the O(B) set-up (trivial initial estimate, priors) is a handful of torch ops here,
everything that touches pixels -- residuals, robust weights, Jacobians, J^T W J, the damped
Cholesky solve, the manifold update, the lambda schedule, early stopping, the covariance -- runs
inside libgclm.so in `num_steps + 1` kernel launches with no host synchronisation.

Documented divergences from the reference (SURVEY.md 8a "quirks"):
  * a failed Cholesky zeroes the step of THAT image only (reference: whole batch, :129-133);
  * with `prior_focal` AND a distortion model the reference indexes the gravity step for the
    distortion update (:224-235, a bug); here the distortion column is used;
  * `stop_mode="per_image"` (opt-in) freezes images individually; the default "batch" reproduces
    the reference's batch-global `torch.allclose` criterion (:90-92).
"""

from __future__ import annotations

import logging
from types import SimpleNamespace
from typing import Any, Dict, Optional, Tuple

import torch
import torch.nn as nn

from . import _lib
from .camera import BaseCamera, camera_models, focal2fov, wrap_camera
from .engine import get_engine
from .gravity import Gravity, wrap_gravity

logger = logging.getLogger(__name__)

UNCERTAINTY_KEYS = ("roll_uncertainty", "pitch_uncertainty", "gravity_uncertainty",
                    "focal_uncertainty", "vfov_uncertainty")


_INIT_CACHE: Dict[tuple, Tuple[torch.Tensor, torch.Tensor]] = {}


def get_trivial_estimation(data: Dict[str, torch.Tensor], camera_model) -> Tuple[torch.Tensor, torch.Tensor]:
    """Initial estimate (reference: lm_optimizer.py:20-58): roll = pitch = 0, f = 0.7 max(h, w)
    passed through a vfov round trip, fx rescaled by `scales`, priors override.

    Returns raw tensors (camera (B,8) float32, gravity (B,3) float32) on the fields' device.
    """
    if "up_field" in data:
        ref = data["up_field"]
    elif "latitude_field" in data:
        ref = data["latitude_field"]
    else:
        ref = data["packed_fields"]["up"]  # pre-transformed fields (Engine.pack_*): shape and device only
    ref = ref.detach()
    B = ref.shape[0]
    h, w = ref.shape[-2:]
    plain = not any(k in data for k in ("prior_focal", "prior_gravity", "prior_dist", "scales"))
    key = (B, h, w, str(ref.device), camera_model.name())
    if plain and key in _INIT_CACHE:  # the prior-free estimate depends on the shape only
        cam, grav = _INIT_CACHE[key]
        return cam.clone(), grav.clone()  # callers own what they get (no aliasing between solves)
    hh, ww = ref.new_full((B,), float(h)), ref.new_full((B,), float(w))
    focal = data["prior_focal"] if "prior_focal" in data else 0.7 * torch.max(hh, ww)
    params = {"width": ww, "height": hh, "vfov": focal2fov(focal.to(ref), h)}
    if "scales" in data:
        params["scales"] = data["scales"].to(ref)
    if "prior_dist" in data:
        params["dist"] = torch.as_tensor(data["prior_dist"]).to(ref)
    cam = camera_model.from_dict(params)._data.float().to(ref.device)
    if cam.dim() == 1:
        cam = cam[None]
    if "prior_gravity" in data:
        g = data["prior_gravity"]
        g = g.vec3d if hasattr(g, "vec3d") else g
        grav = Gravity(g.float().to(ref.device))._data
    else:
        zeros = ref.new_zeros((B,))
        grav = Gravity.from_rp(zeros, zeros)._data.float()
    out = cam.contiguous(), grav.reshape(B, 3).contiguous()
    if plain:
        if len(_INIT_CACHE) > 64:
            _INIT_CACHE.clear()
        _INIT_CACHE[key] = (out[0].clone(), out[1].clone())
    return out


class LMOptimizer(nn.Module):
    """Levenberg-Marquardt calibration on the B200 engine (same surface as the reference)."""

    default_conf = {
        # camera model
        "camera_model": "pinhole",
        "shared_intrinsics": False,
        # LM schedule
        "num_steps": 30,
        "lambda_": 0.1,
        "fix_lambda": False,
        "early_stop": True,
        "atol": 1e-8,
        "rtol": 1e-8,
        "use_spherical_manifold": True,
        "use_log_focal": True,
        # robust loss
        "up_loss_fn_scale": 1e-2,
        "lat_loss_fn_scale": 1e-2,
        "verbose": False,
        # engine-only knobs (not in the reference)
        "stop_mode": "batch",  # "batch" (reference semantics) | "per_image"
        "rows_per_tile": 0,
        "schedule": "auto",  # "auto" | "per_pass" | "persistent" (one cooperative launch; small batches)
        # "float16": the fields are transformed once per solve into the 16 B/pixel layout (sin(latitude) + fp16
        # confidences, gclm_pack_fields) and every pass streams that instead of 20 B/pixel
        "confidence_dtype": "float32",
    }

    def __init__(self, conf: Optional[Dict[str, Any]] = None):
        super().__init__()
        self.conf = conf = SimpleNamespace(**{**self.default_conf, **(conf or {})})
        self.num_steps = conf.num_steps
        self.comm = None  # gclm communicator for multi-GPU shared intrinsics (parallel.py)
        self.set_camera_model(conf.camera_model)
        self.setup_optimization_and_priors(shared_intrinsics=conf.shared_intrinsics)

    @classmethod
    def from_reference(cls, ref_optimizer) -> "LMOptimizer":
        """Build from a live reference `LMOptimizer` (copies its conf and current attributes)."""
        conf = {k: v for k, v in vars(ref_optimizer.conf).items() if k in cls.default_conf}
        new = cls(conf)
        new.num_steps = ref_optimizer.num_steps
        new.set_camera_model(ref_optimizer.camera_model.name())
        new.shared_intrinsics = ref_optimizer.shared_intrinsics
        new.train(ref_optimizer.training)
        return new

    # ---- reference-compatible mutators ---------------------------------------------------------
    def set_camera_model(self, camera_model: str) -> None:
        assert camera_model in camera_models, (
            f"Unknown camera model: {camera_model} not in {camera_models.keys()}"
        )
        self.camera_model = camera_models[camera_model]
        self.camera_has_distortion = self.camera_model.num_dist_params() > 0

    def setup_optimization_and_priors(self, data: Optional[Dict[str, torch.Tensor]] = None,
                                      shared_intrinsics: bool = False) -> None:
        data = data or {}
        self.shared_intrinsics = shared_intrinsics
        self.estimate_gravity = "prior_gravity" not in data
        self.estimate_focal = "prior_focal" not in data
        self.estimate_dist = self.camera_has_distortion and "prior_dist" not in data
        K = self.camera_model.num_dist_params()
        self.gravity_delta_dims = (0, 1) if self.estimate_gravity else (-1,)
        nxt = 2 if self.estimate_gravity else 0
        self.focal_delta_dims = (nxt,) if self.estimate_focal else (-1,)
        nxt += int(self.estimate_focal)
        self.dist_delta_dims = tuple(range(nxt, nxt + K)) if self.estimate_dist else None
        self.n_intrinsic_params = int(self.estimate_focal) + (K if self.camera_has_distortion else 0)

    # ---- engine plumbing -------------------------------------------------------------------------
    def _options(self) -> _lib.Options:
        c = self.conf
        opt = _lib.default_options()
        opt.num_steps = int(self.num_steps)
        opt.lambda_ = float(c.lambda_)
        opt.fix_lambda = int(bool(c.fix_lambda))
        opt.early_stop = int(bool(c.early_stop))
        opt.atol, opt.rtol = float(c.atol), float(c.rtol)
        opt.use_spherical_manifold = int(bool(c.use_spherical_manifold))
        opt.use_log_focal = int(bool(c.use_log_focal))
        opt.up_loss_fn_scale = float(c.up_loss_fn_scale)
        opt.lat_loss_fn_scale = float(c.lat_loss_fn_scale)
        opt.stop_mode = {"batch": _lib.STOP_BATCH, "per_image": _lib.STOP_PER_IMAGE}[c.stop_mode]
        opt.compute_uncertainty = int(not self.training)  # reference: `if not self.training` (:635)
        opt.rows_per_tile = int(c.rows_per_tile)
        opt.schedule = {"auto": _lib.SCHEDULE_AUTO, "per_pass": _lib.SCHEDULE_PER_PASS,
                        "persistent": _lib.SCHEDULE_PERSISTENT}[c.schedule]
        return opt

    @staticmethod
    def _field(data, key, shape_tail):
        t = data.get(key)
        if t is None:
            return None
        t = t.detach()
        if t.dtype != torch.float32:
            t = t.float()
        return t.contiguous()

    def optimize(self, data: Dict[str, torch.Tensor], camera_init: torch.Tensor,
                 gravity_init: torch.Tensor) -> Tuple[BaseCamera, Gravity, Dict[str, torch.Tensor]]:
        """Run the engine from an explicit initial estimate (reference: :551-644).

        `camera_init` / `gravity_init` may be raw tensors or Camera / Gravity objects.
        """
        cam0 = camera_init._data if hasattr(camera_init, "_data") else camera_init
        g0 = gravity_init._data if hasattr(gravity_init, "_data") else gravity_init
        packed = data.get("packed_fields")  # pre-transformed by the caller (Engine.pack_decoder_outputs / pack_fields)
        up = self._field(data, "up_field", 2)
        lat = self._field(data, "latitude_field", 1)
        ref = packed["up"] if packed is not None else (up if up is not None else lat)
        if ref is None:
            raise KeyError("latitude_field")
        has_up, has_lat = packed is not None or up is not None, packed is not None or lat is not None
        if not ref.is_cuda:
            raise _lib.GclmError(
                "geocalib_b200.LMOptimizer runs on CUDA tensors only (fields are on "
                f"{ref.device}); use the reference optimizer for CPU execution"
            )
        uconf = self._field(data, "up_confidence", 0) if up is not None else None
        lconf = self._field(data, "latitude_confidence", 0) if lat is not None else None
        if torch.is_grad_enabled() and any(
                isinstance(data.get(k), torch.Tensor) and data[k].requires_grad
                for k in ("up_field", "latitude_field", "up_confidence", "latitude_confidence")):
            # the engine is inference-only (raw kernels, no backward): cutting gradients silently would
            # turn a training run through the LM (siclib) into a no-op
            raise _lib.GclmError(
                "geocalib_b200.LMOptimizer is inference-only: the fields require grad but the engine has no "
                "backward pass; call it under torch.no_grad() or detach the fields")
        name = self.camera_model.name()
        eng = get_engine(ref.device)
        if packed is None and self.conf.confidence_dtype == "float16" and up is not None and lat is not None \
                and not self.shared_intrinsics and ref.shape[-1] % 4 == 0:
            packed = eng.pack_fields(up, lat, uconf, lconf, conf_dtype=torch.float16)
        raw = eng.solve(
            name, up, lat, uconf, lconf,
            cam0.detach().float().contiguous().to(ref.device), g0.detach().float().contiguous().to(ref.device),
            options=self._options(), estimate_gravity=self.estimate_gravity,
            estimate_focal=self.estimate_focal, estimate_dist=self.estimate_dist,
            shared_intrinsics=bool(self.shared_intrinsics), comm=self.comm, packed=packed,
        )
        costs = raw["costs"]
        infos: Dict[str, torch.Tensor] = {"stop_at": raw["stop_at"].to(torch.float32)}
        if has_up:
            infos["initial_up_cost"] = costs[:, 1]
        if has_lat:
            infos["initial_latitude_cost"] = costs[:, 2]
        infos["initial_cost"] = costs[:, 0]
        if not self.training:
            infos["covariance"] = raw["covariance"]
            for i, k in enumerate(UNCERTAINTY_KEYS):
                infos[k] = raw["uncertainty"][:, i]
        if has_up:
            infos["final_up_cost"] = costs[:, 4]
        if has_lat:
            infos["final_latitude_cost"] = costs[:, 5]
        infos["final_cost"] = costs[:, 3]
        infos["status"] = raw["status"]
        return wrap_camera(name, raw["camera"]), wrap_gravity(raw["gravity"]), infos

    def forward(self, data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """Run the LM optimisation (reference: lm_optimizer.py:646-664)."""
        camera_init, gravity_init = get_trivial_estimation(data, self.camera_model)
        self.setup_optimization_and_priors(data, shared_intrinsics=self.shared_intrinsics)
        camera_opt, gravity_opt, infos = self.optimize(data, camera_init, gravity_init)
        if self.conf.verbose:
            logger.info("focal (px): %s -> %s", camera_init[:, 3].tolist(), camera_opt.f[..., 1].tolist())
        return {"camera": camera_opt, "gravity": gravity_opt, **infos}


B200LMOptimizer = LMOptimizer


def install(model: nn.Module) -> LMOptimizer:
    """Swap the LM optimizer of a reference model for the B200 engine, in place.

    Accepts the reference extractor (`geocalib.GeoCalib`, whose `.model` is the network,
    extractor.py:41-43) or the network itself (`geocalib.geocalib.GeoCalib`, geocalib.py:106).
    No reference file is edited: this is an attribute swap of `model.optimizer`.
    """
    net = model.model if hasattr(model, "model") and hasattr(model.model, "optimizer") else model
    if not hasattr(net, "optimizer"):
        raise AttributeError("expected a GeoCalib model with an `optimizer` attribute")
    net.optimizer = LMOptimizer.from_reference(net.optimizer)
    return net.optimizer
