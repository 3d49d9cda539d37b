"""Camera parameter containers returned by the B200 LM engine.

Host-side mirror of the boundary types of the reference (`geocalib/camera.py:18-152, 522-950`,
`geocalib/misc.py:50-154`): same class names, same `_data` layout (B, 8) =
[w, h, fx, fy, cx, cy, k1, k2], same property / method names, so objects returned by
`geocalib_b200.LMOptimizer` can be used wherever the reference's `BaseCamera` objects are
(`extractor._post_process` calls `undo_scale_crop`, `size`, `f`; `print_calibration` reads
`vfov`, `f`, `dist`).  Only the O(B) bookkeeping lives here -- the per-pixel projection /
distortion math of the reference classes is the CUDA kernel's job (csrc/gclm_device.cuh).

When the reference package itself is importable, `wrap_camera()` returns the reference's own
class instead, which makes the swap invisible to isinstance checks.
"""

from __future__ import annotations

import math
from typing import Dict, Sequence, Union

import torch

Number = Union[int, float]


class TensorStruct:
    """A (..., D) tensor with named slices; batch shape is `shape` (reference: TensorWrapper)."""

    def __init__(self, data: torch.Tensor):
        if not isinstance(data, torch.Tensor):
            data = torch.as_tensor(data)
        self._data = data

    # ---- tensor plumbing -------------------------------------------------------------------
    shape = property(lambda self: self._data.shape[:-1])
    device = property(lambda self: self._data.device)
    dtype = property(lambda self: self._data.dtype)

    def _like(self, data: torch.Tensor):
        return type(self)(data)

    def __getitem__(self, index):
        return self._like(self._data[index])

    def __setitem__(self, index, item):
        self._data[index] = item._data if isinstance(item, TensorStruct) else item

    def __len__(self):
        return self._data.shape[0]

    def to(self, *args, **kwargs):
        return self._like(self._data.to(*args, **kwargs))

    def cpu(self):
        return self._like(self._data.cpu())

    def cuda(self):
        return self._like(self._data.cuda())

    def float(self):
        return self._like(self._data.float())

    def double(self):
        return self._like(self._data.double())

    def detach(self):
        return self._like(self._data.detach())

    def pin_memory(self):
        return self._like(self._data.pin_memory())

    def numpy(self):
        return self._data.detach().cpu().numpy()

    def unsqueeze(self, *a, **k):
        return self._like(self._data.unsqueeze(*a, **k))

    def squeeze(self, *a, **k):
        return self._like(self._data.squeeze(*a, **k))

    def new_tensor(self, *a, **k):
        return self._data.new_tensor(*a, **k)

    def new_zeros(self, *a, **k):
        return self._data.new_zeros(*a, **k)

    def new_ones(self, *a, **k):
        return self._data.new_ones(*a, **k)

    def new_full(self, *a, **k):
        return self._data.new_full(*a, **k)

    def new_empty(self, *a, **k):
        return self._data.new_empty(*a, **k)

    @classmethod
    def stack(cls, objects: Sequence["TensorStruct"], dim: int = 0):
        return cls(torch.stack([o._data for o in objects], dim=dim))

    def __repr__(self):
        return f"{type(self).__name__} {tuple(self.shape)} {self.dtype} {self.device}"


def focal2fov(focal: torch.Tensor, size: torch.Tensor) -> torch.Tensor:
    """Field of view of a pinhole of focal `focal` over `size` pixels (geocalib/utils.py:277)."""
    return 2 * torch.atan(size / (2 * focal))


def fov2focal(fov: torch.Tensor, size: torch.Tensor) -> torch.Tensor:
    """Inverse of focal2fov (geocalib/utils.py:272)."""
    return size / 2 / torch.tan(fov / 2)


class BaseCamera(TensorStruct):
    """[w, h, fx, fy, cx, cy, k1, k2]; shorter inputs (6 or 7 wide) are zero padded to 8."""

    NUM_DIST = 0
    MODEL_NAME = "base"
    DIST_RANGE = (-0.7, 0.7)

    def __init__(self, data: torch.Tensor):
        data = torch.as_tensor(data)
        width = data.shape[-1]
        if width not in (6, 7, 8):
            raise AssertionError(data.shape)
        if width < 8:
            data = torch.cat([data, data.new_zeros(data.shape[:-1] + (8 - width,))], -1)
        super().__init__(data)

    # ---- identity --------------------------------------------------------------------------
    @classmethod
    def name(cls) -> str:
        return cls.MODEL_NAME

    @classmethod
    def num_dist_params(cls) -> int:
        return cls.NUM_DIST

    @classmethod
    def model_id(cls) -> int:
        return {"pinhole": 0, "simple_radial": 1, "radial": 2, "simple_divisional": 3}[cls.MODEL_NAME]

    # ---- construction (reference: BaseCamera.from_dict, camera.py:49-93) ---------------------
    @classmethod
    def from_dict(cls, params: Dict[str, torch.Tensor]) -> "BaseCamera":
        p = {k: (v if isinstance(v, torch.Tensor) else torch.tensor(v)) for k, v in params.items()}
        h, w = p["height"], p["width"]
        if "f" in p:
            f = p["f"]
        elif "vfov" in p:
            f = fov2focal(p["vfov"], h)
        else:
            raise ValueError("Focal length or vertical field of view must be provided.")
        zeros = torch.zeros_like(f)
        if "dist" in p:
            k1 = p["dist"][..., (0,)]
            k2 = p["dist"][..., (1,)] if p["dist"].shape[-1] == 2 else torch.zeros_like(k1)
        elif "k1_hat" in p:
            k1 = p["k1_hat"] * (f / h) ** 2
            k2 = p.get("k2", torch.zeros_like(k1))
        else:
            k1, k2 = p.get("k1", zeros), p.get("k2", zeros)
        fx = f
        if "scales" in p:  # anisotropic resize: only fx changes (camera.py:89-90)
            fx = f * p["scales"][..., 0] / p["scales"][..., 1]
        cols = [w, h, fx, f, p.get("cx", w / 2), p.get("cy", h / 2), k1, k2]
        return cls(torch.stack(cols, dim=-1))

    def pinhole(self):
        return type(self)(self._data[..., :6])

    # ---- named views -----------------------------------------------------------------------
    size = property(lambda self: self._data[..., 0:2])
    f = property(lambda self: self._data[..., 2:4])
    c = property(lambda self: self._data[..., 4:6])
    vfov = property(lambda self: focal2fov(self._data[..., 3], self._data[..., 1]))
    hfov = property(lambda self: focal2fov(self._data[..., 2], self._data[..., 0]))

    @property
    def K(self) -> torch.Tensor:
        """(..., 3, 3) intrinsic matrix."""
        K = self._data.new_zeros(tuple(self.shape) + (3, 3))
        K[..., 0, 0], K[..., 1, 1] = self._data[..., 2], self._data[..., 3]
        K[..., 0, 2], K[..., 1, 2] = self._data[..., 4], self._data[..., 5]
        K[..., 2, 2] = 1
        return K

    def _dist_block(self) -> torch.Tensor:
        return self.dist if self.NUM_DIST else self._data.new_zeros(self.f.shape)

    # ---- image-space transforms (camera.py:154-169, 206-209) ----------------------------------
    def scale(self, scales):
        s = (scales, scales) if isinstance(scales, (int, float)) else scales
        s = s if isinstance(s, torch.Tensor) else self.new_tensor(s)
        return type(self)(torch.cat([self.size * s, self.f * s, self.c * s, self._dist_block()], -1))

    def crop(self, pad):
        pad = pad if isinstance(pad, torch.Tensor) else self.new_tensor(pad)
        pad = pad.to(self._data)
        return type(self)(torch.cat([self.size + pad, self.f, self.c + pad / 2, self._dist_block()], -1))

    def undo_scale_crop(self, data: Dict[str, torch.Tensor]):
        cam = self.crop(-data["crop_pad"]) if "crop_pad" in data else self
        return cam.scale(1.0 / data["scales"])

    # ---- O(B) parameter updates (host versions of csrc focal_update / dist_update) ------------
    def update_focal(self, delta: torch.Tensor, as_log: bool = False):
        """camera.py:136-152: move fy (log or linear), clamp vfov to [5, 150] deg, keep fx/fy."""
        f_old = self.f
        f_new = torch.exp(torch.log(f_old) + delta) if as_log else f_old + delta
        h = self._data[..., 1]
        lo = fov2focal(torch.full_like(h, math.radians(150.0)), h)
        hi = fov2focal(torch.full_like(h, math.radians(5.0)), h)
        fy = torch.minimum(torch.maximum(f_new[..., 1], lo), hi)
        fx = fy * f_old[..., 0] / f_old[..., 1]
        return type(self)(torch.cat([self.size, torch.stack([fx, fy], -1), self.c, self._dist_block()], -1))

    def update_dist(self, delta: torch.Tensor, dist_range=None):
        """camera.py:599-604 etc.: k <- clamp(k + delta); delta broadcasts over the dist block."""
        lo, hi = dist_range or self.DIST_RANGE
        dist = (self.dist + self.new_ones(self.dist.shape) * delta).clamp(lo, hi)
        return type(self)(torch.cat([self.size, self.f, self.c, dist], -1))


class Pinhole(BaseCamera):
    """No distortion (reference: camera.py:522-562)."""

    MODEL_NAME = "pinhole"


class SimpleRadial(BaseCamera):
    """1 + k1 r^2 (reference: camera.py:565-660).  `dist` is the 2-wide tail of `_data`."""

    MODEL_NAME, NUM_DIST = "simple_radial", 1
    dist = property(lambda self: self._data[..., 6:])
    k1 = property(lambda self: self._data[..., 6])
    k1_hat = property(lambda self: self._data[..., 6] / (self._data[..., 3] / self._data[..., 1]) ** 2)


class Radial(BaseCamera):
    """1 + k1 r^2 + k2 r^4 (reference: camera.py:663-786)."""

    MODEL_NAME, NUM_DIST = "radial", 2
    dist = property(lambda self: self._data[..., 6:8])
    k1 = property(lambda self: self._data[..., 6])
    k2 = property(lambda self: self._data[..., 7])


class SimpleDivisional(BaseCamera):
    """(1 - sqrt(1 - 4 k1 r^2)) / (2 k1 r^2) (reference: camera.py:789-942)."""

    MODEL_NAME, NUM_DIST, DIST_RANGE = "simple_divisional", 1, (-3.0, 3.0)
    dist = property(lambda self: self._data[..., 6:])
    k1 = property(lambda self: self._data[..., 6])


camera_models = {
    "pinhole": Pinhole,
    "radial": Radial,
    "simple_radial": SimpleRadial,
    "simple_divisional": SimpleDivisional,
}


def _reference_module(name: str):
    """The reference package, but only if the host application already imported it."""
    import sys

    return sys.modules.get(name)


def wrap_camera(model: str, data: torch.Tensor):
    """Camera object for `data` (B, 8): the reference's class if it is loaded, else ours."""
    ref = _reference_module("geocalib.camera")
    if ref is not None and hasattr(ref, "camera_models"):
        return ref.camera_models[model](data)
    return camera_models[model](data)
