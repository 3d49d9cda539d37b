"""Thin Python handle over the C ABI: device tensors in, device tensors out.

torch is used for device memory and streams only; all arithmetic on the fields happens inside
libgclm.so.  Inputs must already be CUDA float32 tensors -- there is no CPU path.
"""

from __future__ import annotations

import ctypes
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import CAMERA_MODEL_IDS, Options, PackedFields, Problem, Result

NUM_DIST = {"pinhole": 0, "simple_radial": 1, "radial": 2, "simple_divisional": 1}

_engines: Dict[int, "Engine"] = {}


def get_engine(device: Optional[torch.device] = None) -> "Engine":
    """One engine (scratch allocator) per CUDA device, created on first use."""
    if not torch.cuda.is_available():
        raise _lib.GclmError("geocalib_b200 needs a CUDA device (sm_100a); none is visible")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    if idx not in _engines:
        _engines[idx] = Engine(idx)
    return _engines[idx]


def _dev_ptr(t: Optional[torch.Tensor], shape, name: str, device) -> Optional[int]:
    if t is None:
        return None
    if not t.is_cuda or t.device != device:
        raise _lib.GclmError(f"{name} must live on {device} (got {t.device}); there is no CPU path")
    if t.dtype != torch.float32:
        raise _lib.GclmError(f"{name} must be float32 (got {t.dtype})")
    if tuple(t.shape) != tuple(shape):
        raise _lib.GclmError(f"{name} has shape {tuple(t.shape)}, expected {tuple(shape)}")
    if not t.is_contiguous():
        raise _lib.GclmError(f"{name} must be contiguous")
    return t.data_ptr()


class Engine:
    """Owns a gclm_handle on one device."""

    def __init__(self, device_index: int):
        self.lib = _lib.load()
        self.device = torch.device("cuda", device_index)
        h = ctypes.c_void_p()
        _lib.check(self.lib.gclm_create(device_index, ctypes.byref(h)), "gclm_create")
        self._h = h

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                self.lib.gclm_destroy(self._h)
                self._h = None
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.gclm_launch_count(self._h))

    @property
    def persistent_count(self) -> int:
        """Solves that ran as ONE cooperative launch (the small-batch schedule)."""
        return int(self.lib.gclm_persistent_count(self._h))

    def pass_timing(self, enable: bool) -> None:
        """Start (and reset) / stop the opt-in event timing of the pass launches (bench.py's roofline)."""
        _lib.check(self.lib.gclm_pass_timing(self._h, int(bool(enable))), "gclm_pass_timing")

    def pass_time(self):
        """(total ms, number of pass launches) of the solves enqueued since pass_timing(True); synchronises."""
        ms, n = ctypes.c_double(0.0), ctypes.c_int64(0)
        _lib.check(self.lib.gclm_pass_time(self._h, ctypes.byref(ms), ctypes.byref(n)), "gclm_pass_time")
        return float(ms.value), int(n.value)

    def reserve(self, batch: int, height: int, width: int) -> None:
        _lib.check(self.lib.gclm_reserve(self._h, batch, height, width), "gclm_reserve")

    # ------------------------------------------------------------------------------------------
    def solve(
        self,
        camera_model: str,
        up_field: Optional[torch.Tensor],
        latitude_field: Optional[torch.Tensor],
        up_confidence: Optional[torch.Tensor],
        latitude_confidence: Optional[torch.Tensor],
        camera_init: torch.Tensor,
        gravity_init: torch.Tensor,
        *,
        options: Optional[Options] = None,
        estimate_gravity: bool = True,
        estimate_focal: bool = True,
        estimate_dist: bool = True,
        shared_intrinsics: bool = False,
        comm=None,
        out: Optional[Dict[str, torch.Tensor]] = None,
        packed: Optional[Dict[str, torch.Tensor]] = None,
    ) -> Dict[str, torch.Tensor]:
        """Enqueue one batched LM solve on the current stream; returns raw result tensors.

        `packed`: pre-transformed fields from pack_fields / pack_decoder_outputs ({"up", "sin_lat", "up_conf",
        "lat_conf"}); the four field arguments are then ignored (gclm_solve_packed)."""
        if packed is not None:
            up_field, latitude_field, up_confidence, latitude_confidence = packed["up"], None, None, None
        ref = up_field if up_field is not None else latitude_field
        if ref is None:
            raise _lib.GclmError("need at least one of up_field / latitude_field")
        dev = ref.device
        B, _, H, W = ref.shape
        opt = options if options is not None else _lib.default_options()
        n = 2 * int(estimate_gravity) + int(estimate_focal) + NUM_DIST[camera_model]
        prob = Problem(
            batch=B, height=H, width=W, camera_model=CAMERA_MODEL_IDS[camera_model],
            up_field=_dev_ptr(up_field, (B, 2, H, W), "up_field", dev),
            latitude_field=_dev_ptr(latitude_field, (B, 1, H, W), "latitude_field", dev),
            up_confidence=_dev_ptr(up_confidence, (B, H, W), "up_confidence", dev),
            latitude_confidence=_dev_ptr(latitude_confidence, (B, H, W), "latitude_confidence", dev),
            camera_init=_dev_ptr(camera_init, (B, 8), "camera_init", dev),
            gravity_init=_dev_ptr(gravity_init, (B, 3), "gravity_init", dev),
            estimate_gravity=int(estimate_gravity), estimate_focal=int(estimate_focal),
            estimate_dist=int(estimate_dist), shared_intrinsics=int(shared_intrinsics),
        )
        if out is None:
            f32 = dict(device=dev, dtype=torch.float32)
            out = {
                "camera": torch.empty((B, 8), **f32),
                "gravity": torch.empty((B, 3), **f32),
                "covariance": torch.zeros((B, n, n), **f32),
                "uncertainty": torch.zeros((B, 5), **f32),
                "costs": torch.empty((B, 6), **f32),
                "stop_at": torch.empty((B,), device=dev, dtype=torch.int32),
                "status": torch.empty((B,), device=dev, dtype=torch.int32),
                "hessian_rpf": torch.empty((B, n, n), **f32),
            }
        res = Result(**{k: out[k].data_ptr() for k in
                        ("camera", "gravity", "covariance", "uncertainty", "costs", "stop_at", "status", "hessian_rpf")})
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        # the library runs on its handle's device and restores the caller's current device (DeviceGuard)
        if packed is not None:
            if shared_intrinsics:
                raise _lib.GclmError("packed fields: per-image intrinsics only")
            f16 = packed["up_conf"].dtype == torch.float16
            for k in ("sin_lat", "up_conf", "lat_conf"):
                t = packed[k]
                if not t.is_cuda or t.device != dev or not t.is_contiguous() or t.numel() != B * H * W:
                    raise _lib.GclmError(f"packed[{k!r}]: expected a contiguous CUDA tensor with {B * H * W} elements")
            pf = PackedFields(batch=B, height=H, width=W, conf_format=_lib.CONF_F16 if f16 else _lib.CONF_F32,
                              up=up_field.data_ptr(), sin_lat=packed["sin_lat"].data_ptr(),
                              up_conf=packed["up_conf"].data_ptr(), lat_conf=packed["lat_conf"].data_ptr())
            prob.up_field = None
            st = self.lib.gclm_solve_packed(self._h, ctypes.byref(pf), ctypes.byref(prob), ctypes.byref(opt),
                                            ctypes.byref(res), stream)
            _lib.check(st, "gclm_solve_packed")
        elif shared_intrinsics:
            comm_h = getattr(comm, "handle", comm)
            st = self.lib.gclm_solve_shared(self._h, ctypes.byref(prob), ctypes.byref(opt),
                                            ctypes.byref(res), comm_h, stream)
            _lib.check(st, "gclm_solve_shared")
        else:
            st = self.lib.gclm_solve(self._h, ctypes.byref(prob), ctypes.byref(opt),
                                     ctypes.byref(res), stream)
            _lib.check(st, "gclm_solve")
        return out

    # ------------------------------------------------------------------------------------------
    def pack_fields(self, up_field: torch.Tensor, latitude_field: torch.Tensor,
                    up_confidence: Optional[torch.Tensor], latitude_confidence: Optional[torch.Tensor], *,
                    conf_dtype: torch.dtype = torch.float16) -> Dict[str, torch.Tensor]:
        """gclm_pack_fields: the five activated planes -> the layout the engine streams:
        {"up" (as given), "sin_lat" fp32, "up_conf" / "lat_conf" in `conf_dtype` (float16: 16 B/pixel/pass)}."""
        dev = up_field.device
        B, _, H, W = up_field.shape
        prob = Problem(
            batch=B, height=H, width=W, camera_model=0,
            up_field=_dev_ptr(up_field, (B, 2, H, W), "up_field", dev),
            latitude_field=_dev_ptr(latitude_field, (B, 1, H, W), "latitude_field", dev),
            up_confidence=_dev_ptr(up_confidence, (B, H, W), "up_confidence", dev),
            latitude_confidence=_dev_ptr(latitude_confidence, (B, H, W), "latitude_confidence", dev))
        assert conf_dtype in (torch.float16, torch.float32)
        sin_lat = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
        uc = torch.empty((B, H, W), device=dev, dtype=conf_dtype)
        lc = torch.empty((B, H, W), device=dev, dtype=conf_dtype)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self.lib.gclm_pack_fields(self._h, ctypes.byref(prob),
                                             _lib.CONF_F16 if conf_dtype == torch.float16 else _lib.CONF_F32,
                                             sin_lat.data_ptr(), uc.data_ptr(), lc.data_ptr(), stream), "gclm_pack_fields")
        return {"up": up_field, "sin_lat": sin_lat, "up_conf": uc, "lat_conf": lc}

    def pack_decoder_outputs(self, up_logits: torch.Tensor, latitude_logits: torch.Tensor,
                             up_conf_logits: torch.Tensor, lat_conf_logits: torch.Tensor, *,
                             conf_dtype: torch.dtype = torch.float16) -> Dict[str, torch.Tensor]:
        """gclm_pack_decoder_outputs: the fused decoder tail (geocalib.py:44-89) -- from the heads'
        PRE-activation outputs straight to the streamed layout (F.normalize, clamp(tanh), sigmoid)."""
        dev = up_logits.device
        B, _, H, W = up_logits.shape
        ptrs = [_dev_ptr(up_logits, (B, 2, H, W), "up_logits", dev),
                _dev_ptr(latitude_logits, (B, 1, H, W), "latitude_logits", dev),
                _dev_ptr(up_conf_logits, (B, H, W), "up_conf_logits", dev),
                _dev_ptr(lat_conf_logits, (B, H, W), "lat_conf_logits", dev)]
        up = torch.empty((B, 2, H, W), device=dev, dtype=torch.float32)
        sin_lat = torch.empty((B, 1, H, W), device=dev, dtype=torch.float32)
        uc = torch.empty((B, H, W), device=dev, dtype=conf_dtype)
        lc = torch.empty((B, H, W), device=dev, dtype=conf_dtype)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self.lib.gclm_pack_decoder_outputs(
            self._h, B, H, W, *ptrs, _lib.CONF_F16 if conf_dtype == torch.float16 else _lib.CONF_F32,
            up.data_ptr(), sin_lat.data_ptr(), uc.data_ptr(), lc.data_ptr(), stream), "gclm_pack_decoder_outputs")
        return {"up": up, "sin_lat": sin_lat, "up_conf": uc, "lat_conf": lc}

    # ------------------------------------------------------------------------------------------
    def system(self, camera_model: str, up_field, latitude_field, up_confidence, latitude_confidence,
               camera: torch.Tensor, gravity: torch.Tensor, *, options: Optional[Options] = None,
               rpf_basis: bool = False):
        """gclm_system: (G (B,NF), H (B,NF,NF), costs (B,2)) of one pass at the given parameters
        (reference: LMOptimizer.setup_system + calculate_costs, lm_optimizer.py:276-315,387-461)."""
        ref = up_field if up_field is not None else latitude_field
        dev = ref.device
        B, _, H, W = ref.shape
        opt = options if options is not None else _lib.default_options()
        nf = 3 + NUM_DIST[camera_model]
        prob = Problem(
            batch=B, height=H, width=W, camera_model=CAMERA_MODEL_IDS[camera_model],
            up_field=_dev_ptr(up_field, (B, 2, H, W), "up_field", dev),
            latitude_field=_dev_ptr(latitude_field, (B, 1, H, W), "latitude_field", dev),
            up_confidence=_dev_ptr(up_confidence, (B, H, W), "up_confidence", dev),
            latitude_confidence=_dev_ptr(latitude_confidence, (B, H, W), "latitude_confidence", dev),
            camera_init=_dev_ptr(camera, (B, 8), "camera", dev),
            gravity_init=_dev_ptr(gravity, (B, 3), "gravity", dev),
            estimate_gravity=1, estimate_focal=1, estimate_dist=1, shared_intrinsics=0,
        )
        f32 = dict(device=dev, dtype=torch.float32)
        G, Hm, c = torch.zeros((B, nf), **f32), torch.zeros((B, nf, nf), **f32), torch.zeros((B, 2), **f32)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self.lib.gclm_system(self._h, ctypes.byref(prob), ctypes.byref(opt), int(bool(rpf_basis)),
                                        G.data_ptr(), Hm.data_ptr(), c.data_ptr(), stream), "gclm_system")
        return G, Hm, c

    # ------------------------------------------------------------------------------------------
    def adjoint(self, camera_model: str, up_field, latitude_field, up_confidence, latitude_confidence,
                camera: torch.Tensor, gravity: torch.Tensor, lam: torch.Tensor, *,
                options: Optional[Options] = None):
        """gclm_adjoint: d(lambda . G)/d(fields) per pixel at (camera, gravity) -- the streaming pass of the
        implicit-function backward (geocalib_b200/autograd.py).  Returns 4 tensors shaped like the inputs."""
        dev = up_field.device
        B, _, H, W = up_field.shape
        opt = options if options is not None else _lib.default_options()
        nf = 3 + NUM_DIST[camera_model]
        prob = Problem(
            batch=B, height=H, width=W, camera_model=CAMERA_MODEL_IDS[camera_model],
            up_field=_dev_ptr(up_field, (B, 2, H, W), "up_field", dev),
            latitude_field=_dev_ptr(latitude_field, (B, 1, H, W), "latitude_field", dev),
            up_confidence=_dev_ptr(up_confidence, (B, H, W), "up_confidence", dev),
            latitude_confidence=_dev_ptr(latitude_confidence, (B, H, W), "latitude_confidence", dev))
        g_up, g_lat = torch.empty_like(up_field), torch.empty_like(latitude_field)
        g_uc = torch.empty((B, H, W), device=dev, dtype=torch.float32)
        g_lc = torch.empty((B, H, W), device=dev, dtype=torch.float32)
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(self.lib.gclm_adjoint(
            self._h, ctypes.byref(prob), ctypes.byref(opt), _dev_ptr(camera.contiguous(), (B, 8), "camera", dev),
            _dev_ptr(gravity.contiguous(), (B, 3), "gravity", dev), _dev_ptr(lam.contiguous(), (B, nf), "lambda", dev),
            g_up.data_ptr(), g_lat.data_ptr(), g_uc.data_ptr(), g_lc.data_ptr(), stream), "gclm_adjoint")
        return g_up, g_lat, g_uc, g_lc

    # ------------------------------------------------------------------------------------------
    def render_fields(self, camera_model: str, camera: torch.Tensor, gravity: torch.Tensor,
                      height: int, width: int, *, sigma: float = 0.0, conf_lo: float = 1.0,
                      seed: int = 0, with_confidence: bool = True) -> Dict[str, torch.Tensor]:
        """Synthetic predicted fields on device (bench / tests): GT render + noise + confidences."""
        dev = camera.device
        B = camera.shape[0]
        f32 = dict(device=dev, dtype=torch.float32)
        up = torch.empty((B, 2, height, width), **f32)
        lat = torch.empty((B, 1, height, width), **f32)
        uc = torch.empty((B, height, width), **f32) if with_confidence else None
        lc = torch.empty((B, height, width), **f32) if with_confidence else None
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        st = self.lib.gclm_render_fields(
            self._h, B, height, width, CAMERA_MODEL_IDS[camera_model],
            _dev_ptr(camera.contiguous(), (B, 8), "camera", dev),
            _dev_ptr(gravity.contiguous(), (B, 3), "gravity", dev),
            float(sigma), float(conf_lo), int(seed) & (2**64 - 1), up.data_ptr(), lat.data_ptr(),
            uc.data_ptr() if uc is not None else None, lc.data_ptr() if lc is not None else None, stream)
        _lib.check(st, "gclm_render_fields")
        out = {"up_field": up, "latitude_field": lat}
        if with_confidence:
            out |= {"up_confidence": uc, "latitude_confidence": lc}
        return out

    # ------------------------------------------------------------------------------------------
    def solve_host(self, camera_model: str, host_fields: Dict[str, torch.Tensor],
                   camera_init: torch.Tensor, gravity_init: torch.Tensor, *,
                   options: Optional[Options] = None, estimate_gravity: bool = True,
                   estimate_focal: bool = True, estimate_dist: bool = True, chunk_images: int = 128,
                   out: Optional[Dict[str, torch.Tensor]] = None) -> Dict[str, torch.Tensor]:
        """gclm_solve_host: HOST (ideally pinned) float32 tensors in, host tensors out; the library
        pipelines H2D / solve / D2H over chunks.  Synchronous."""
        ref = host_fields.get("up_field", host_fields.get("latitude_field"))
        if ref is None:
            raise _lib.GclmError("need at least one of up_field / latitude_field")
        B, _, H, W = ref.shape
        opt = options if options is not None else _lib.default_options()
        n = 2 * int(estimate_gravity) + int(estimate_focal) + NUM_DIST[camera_model]

        def hp(t, shape, name):
            if t is None:
                return None
            if t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous() or tuple(t.shape) != tuple(shape):
                raise _lib.GclmError(f"{name}: expected a contiguous float32 CPU tensor of shape {tuple(shape)}")
            return t.data_ptr()

        prob = Problem(
            batch=B, height=H, width=W, camera_model=CAMERA_MODEL_IDS[camera_model],
            up_field=hp(host_fields.get("up_field"), (B, 2, H, W), "up_field"),
            latitude_field=hp(host_fields.get("latitude_field"), (B, 1, H, W), "latitude_field"),
            up_confidence=hp(host_fields.get("up_confidence"), (B, H, W), "up_confidence"),
            latitude_confidence=hp(host_fields.get("latitude_confidence"), (B, H, W), "latitude_confidence"),
            camera_init=hp(camera_init, (B, 8), "camera_init"), gravity_init=hp(gravity_init, (B, 3), "gravity_init"),
            estimate_gravity=int(estimate_gravity), estimate_focal=int(estimate_focal),
            estimate_dist=int(estimate_dist), shared_intrinsics=0,
        )
        if out is None:
            pin = lambda *shape, dtype=torch.float32: torch.empty(shape, dtype=dtype).pin_memory()
            out = {
                "camera": pin(B, 8), "gravity": pin(B, 3), "covariance": pin(B, n, n), "uncertainty": pin(B, 5),
                "costs": pin(B, 6), "stop_at": pin(B, dtype=torch.int32), "status": pin(B, dtype=torch.int32),
                "hessian_rpf": pin(B, n, n),
            }
        res = Result(**{k: out[k].data_ptr() for k in
                        ("camera", "gravity", "covariance", "uncertainty", "costs", "stop_at", "status", "hessian_rpf")})
        _lib.check(self.lib.gclm_solve_host(self._h, ctypes.byref(prob), ctypes.byref(opt),
                                            ctypes.byref(res), int(chunk_images)), "gclm_solve_host")
        return out
