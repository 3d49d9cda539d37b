"""End-to-end path for HOST-resident fields: pinned host memory -> H2D -> LM solve -> D2H.

The batch is cut into chunks; the host-to-device copy of chunk i+1 runs on a copy stream while
chunk i is being solved on the compute stream (two device staging buffers), so the step time
approaches max(PCIe time, solve time) instead of their sum.  torch supplies the pinned/device
memory and the streams; all arithmetic is the engine's.
"""

from __future__ import annotations

from typing import Dict

import torch

FIELD_KEYS = ("up_field", "latitude_field", "up_confidence", "latitude_confidence")
RESULT_KEYS = ("camera", "gravity", "covariance", "uncertainty", "costs", "stop_at", "status")


class HostPipeline:
    def __init__(self, optimizer, batch: int, height: int, width: int, chunk: int, device: torch.device):
        self.opt, self.B, self.h, self.w, self.dev = optimizer, batch, height, width, device
        self.chunk = max(1, min(chunk, batch))
        shapes = {"up_field": (2, height, width), "latitude_field": (1, height, width),
                  "up_confidence": (height, width), "latitude_confidence": (height, width)}
        self.stage = [
            {k: torch.empty((self.chunk,) + s, device=device, dtype=torch.float32) for k, s in shapes.items()}
            for _ in range(2)
        ]
        self.copy_stream = torch.cuda.Stream(device)
        self.ready = [torch.cuda.Event() for _ in range(2)]   # H2D of buffer i finished
        self.free = [torch.cuda.Event() for _ in range(2)]    # solve reading buffer i finished
        self.h2d_bytes = batch * height * width * 5 * 4
        self.d2h_bytes = 0
        self._host_out = None

    def _alloc_out(self, sample: Dict[str, torch.Tensor]):
        self._host_out = {
            k: torch.empty((self.B,) + tuple(v.shape[1:]), dtype=v.dtype).pin_memory() for k, v in sample.items()
        }
        self.d2h_bytes = sum(v.numel() * v.element_size() for v in self._host_out.values())

    @torch.no_grad()
    def run(self, host_fields: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        """host_fields: pinned CPU tensors with the reference's keys/shapes.  Returns CPU tensors."""
        compute = torch.cuda.current_stream(self.dev)
        keys = [k for k in FIELD_KEYS if k in host_fields]
        starts = list(range(0, self.B, self.chunk))
        for i, s in enumerate(starts):
            e = min(self.B, s + self.chunk)
            buf = self.stage[i % 2]
            with torch.cuda.stream(self.copy_stream):
                if i >= 2:
                    self.copy_stream.wait_event(self.free[i % 2])
                for k in keys:
                    buf[k][: e - s].copy_(host_fields[k][s:e], non_blocking=True)
                self.ready[i % 2].record(self.copy_stream)
            compute.wait_event(self.ready[i % 2])
            data = {k: buf[k][: e - s] for k in keys}
            out = self.opt(data)
            self.free[i % 2].record(compute)
            flat = {
                "camera": out["camera"]._data, "gravity": out["gravity"]._data,
                "costs": torch.stack([out["initial_cost"], out["final_cost"]], -1),
                "stop_at": out["stop_at"], "status": out["status"],
            }
            if "covariance" in out:
                flat["covariance"] = out["covariance"]
                flat["uncertainty"] = torch.stack([out[k] for k in (
                    "roll_uncertainty", "pitch_uncertainty", "gravity_uncertainty",
                    "focal_uncertainty", "vfov_uncertainty")], -1)
            if self._host_out is None:
                self._alloc_out(flat)
            for k, v in flat.items():
                self._host_out[k][s:e].copy_(v, non_blocking=True)
        compute.synchronize()
        return self._host_out
