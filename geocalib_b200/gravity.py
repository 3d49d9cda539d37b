"""Gravity direction container (host-side mirror of `geocalib/gravity.py:12-131`).

A unit 3-vector per image in the camera frame, g = (-sin r cos p, -cos r cos p, sin p) for roll r
and pitch p.  The roll read-back keeps the reference's 1e-4 epsilon (gravity.py:15,62-67) so that
`.rp` of our result and of the reference's agree bit for bit when the vectors do.
"""

from __future__ import annotations

import torch
from torch.nn import functional as F

from .camera import TensorStruct, _reference_module


class Gravity(TensorStruct):
    eps = 1e-4

    def __init__(self, data: torch.Tensor):
        data = torch.as_tensor(data)
        if data.shape[-1] != 3:
            raise AssertionError(data.shape)
        super().__init__(F.normalize(data, dim=-1))

    @classmethod
    def from_rp(cls, roll, pitch) -> "Gravity":
        roll, pitch = torch.as_tensor(roll), torch.as_tensor(pitch)
        cp = torch.cos(pitch)
        return cls(torch.stack([-torch.sin(roll) * cp, -torch.cos(roll) * cp, torch.sin(pitch)], -1))

    vec3d = property(lambda self: self._data)
    x = property(lambda self: self._data[..., 0])
    y = property(lambda self: self._data[..., 1])
    z = property(lambda self: self._data[..., 2])

    @property
    def roll(self) -> torch.Tensor:
        r = torch.asin(-self.x / (torch.sqrt(1 - self.z**2) + self.eps))
        return torch.where(self.y < 0, r, -r - torch.pi * torch.sign(self.x))

    @property
    def pitch(self) -> torch.Tensor:
        return torch.asin(self.z)

    @property
    def rp(self) -> torch.Tensor:
        return torch.stack([self.roll, self.pitch], -1)

    def J_rp(self) -> torch.Tensor:
        """d g / d (roll, pitch), shape (..., 3, 2) (gravity.py:69-101)."""
        r, p = self.roll, self.pitch
        sr, cr, sp, cp = torch.sin(r), torch.cos(r), torch.sin(p), torch.cos(p)
        zero = torch.zeros_like(r)
        d_roll = torch.stack([-cr * cp, sr * cp, zero], -1)
        d_pitch = torch.stack([sr * sp, cr * sp, cp], -1)
        return torch.stack([d_roll, d_pitch], -1)


def wrap_gravity(data: torch.Tensor):
    """Gravity object for `data` (B, 3): the reference's class if it is loaded, else ours."""
    ref = _reference_module("geocalib.gravity")
    if ref is not None and hasattr(ref, "Gravity"):
        return ref.Gravity(data)
    return Gravity(data)
