"""End-to-end path for HOST-resident fields: pinned host memory -> H2D -> LM solve -> D2H.

`HostPipeline.run` is the call a host application without device tensors makes: it computes the
O(B) initial estimate on the CPU (same code as the device path) and hands HOST pointers to
`gclm_solve_host` (include/gclm.h), which cuts the batch into chunks and overlaps the H2D copy of
chunk i+1 with the solve of chunk i on two internal streams (two device staging buffers), so the
step time approaches max(PCIe time, solve time) instead of their sum.
"""

from __future__ import annotations

from typing import Dict

import torch

from .engine import get_engine
from .lm_optimizer import get_trivial_estimation

FIELD_KEYS = ("up_field", "latitude_field", "up_confidence", "latitude_confidence")


class HostPipeline:
    def __init__(self, optimizer, batch: int, height: int, width: int, chunk: int, device: torch.device):
        self.opt, self.B, self.h, self.w, self.dev = optimizer, batch, height, width, device
        self.chunk = max(1, min(chunk, batch))
        self.eng = get_engine(device)
        self.h2d_bytes = batch * (height * width * 5 + 11) * 4
        self.d2h_bytes = 0
        self._out = None

    @torch.no_grad()
    def run(self, host_fields: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """host_fields: (pinned) CPU float32 tensors with the reference's keys/shapes (+ optional
        priors / scales).  Returns pinned CPU tensors (camera (B,8), gravity (B,3), ...)."""
        opt = self.opt
        cam0, g0 = get_trivial_estimation(host_fields, opt.camera_model)
        opt.setup_optimization_and_priors(host_fields, shared_intrinsics=False)
        fields = {k: host_fields[k] for k in FIELD_KEYS if k in host_fields}
        self._out = self.eng.solve_host(
            opt.camera_model.name(), fields, cam0.contiguous(), g0.contiguous(), options=opt._options(),
            estimate_gravity=opt.estimate_gravity, estimate_focal=opt.estimate_focal,
            estimate_dist=opt.estimate_dist, chunk_images=self.chunk, out=self._out)
        if not self.d2h_bytes:
            self.d2h_bytes = sum(v.numel() * v.element_size() for v in self._out.values())
        return self._out
