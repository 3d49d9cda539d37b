"""CUDA-graph replay of the reference CNN in front of the engine (first step).

Once the LM costs ~150 us per image, `GeoCalib.calibrate()` is the CNN: MSCAN + two LightHam decoders issue several
hundred small kernels per image (depthwise strip convolutions `modules.py:374-411`, the NMF2D `bmm` loop `:120-199`)
and at B == 1 every one of them is launch-bound.  `GraphedFrontEnd` captures the UNMODIFIED modules of the reference
network -- `backbone`, `ll_enc`, `perspective_decoder` (geocalib.py:106-110) -- into one CUDA graph per input shape
and replays it from a static input buffer; `install_frontend(model)` puts it in front of the optimizer exactly where
`GeoCalib.forward` (geocalib.py:106-119) calls the three modules.  Weights, arithmetic (fp32, cuDNN / cuBLAS kernels
chosen by torch) and outputs are those of the reference; only the launch mechanism changes.

Two statements of the reference cannot be captured as they are, both because they copy pageable host memory to the
device inside the forward:
  * `NMF2D._build_bases` (modules.py:131-133) draws the random initial bases with the CPU generator and copies them
    over on every forward.  The front end swaps that one method for the same draw (uniform [0,1), normalised over D)
    from the CUDA generator, which torch advances correctly across graph replays; the distribution is the
    reference's, the stream of random numbers is not (it is not reproducible call to call in the reference either);
  * `MSCAN.forward` (modules.py:560) permutes RGB -> BGR by indexing with the Python list [2, 1, 0], which torch turns
    into a CPU index tensor + copy.  The backbone is therefore handed the image inside a proxy that answers exactly
    that indexing expression with the equivalent `flip(1)` (bit-identical values) and forwards everything else.

Not done here (the rest of 8f.4): hand-written kernels for the strip convolutions / NMF loop, channels-last, bf16.
"""

from __future__ import annotations

import types
from typing import Callable, Dict, Optional, Tuple

import torch
import torch.nn.functional as F

FIELD_KEYS = ("up_field", "up_confidence", "latitude_field", "latitude_confidence")
PASS_THROUGH = ("image", "scales", "prior_gravity", "prior_focal", "prior_dist")  # geocalib.py:112-116


def _device_bases(self, B: int, S: int, D: int, R: int, device="cpu") -> torch.Tensor:
    """modules.py:131-133 with the draw made on `device` (capturable; same distribution)."""
    return F.normalize(torch.rand((B * S, D, R), device=device), dim=1)


def nmf_modules(net: torch.nn.Module):
    return [m for m in net.modules() if type(m).__name__ == "NMF2D" and hasattr(m, "_build_bases")]


def use_device_bases(net: torch.nn.Module, build: Optional[Callable] = None) -> Callable[[], None]:
    """Swap `_build_bases` of every NMF2D in `net` (instance attribute; the class is untouched).  Returns the undo."""
    mods = nmf_modules(net)
    for m in mods:
        m._build_bases = types.MethodType(build or _device_bases, m)

    def undo():
        for m in mods:
            m.__dict__.pop("_build_bases", None)

    return undo


class _ChannelFlipProxy:
    """`proxy[:, [2, 1, 0], :, :]` -> `tensor.flip(1)` (modules.py:560 without the host-side index tensor)."""

    def __init__(self, t: torch.Tensor):
        assert t.dim() == 4 and t.shape[1] == 3
        self._t = t

    def __getitem__(self, key):
        full = slice(None)
        if isinstance(key, tuple) and len(key) == 4 and key[1] == [2, 1, 0] and key[0] == full and key[2] == full and key[3] == full:
            return self._t.flip(1)
        return self._t[key]

    def __getattr__(self, name):
        return getattr(self._t, name)


def cnn_forward(net: torch.nn.Module, image: torch.Tensor, capturable: bool = False) -> Dict[str, torch.Tensor]:
    """The CNN part of `GeoCalib.forward` (geocalib.py:106-110): image (B,3,H,W) -> the four decoder outputs.
    `capturable`: hand the backbone the image through `_ChannelFlipProxy` (same values, no host-side index tensor)."""
    data = {"image": image}
    hl = net.backbone({"image": _ChannelFlipProxy(image)} if capturable else data)["features"]
    features = {"hl": hl, "ll": net.ll_enc(data)["features"]}
    return net.perspective_decoder({"features": features})


class GraphedFrontEnd:
    """image -> {up_field, up_confidence, latitude_field, latitude_confidence} as ONE graph launch per call.

    A graph (and its static input / output buffers) is built per (B, H, W) on first use: 3 eager warm-up runs on a side
    stream (cuDNN / cuBLAS pick their algorithms and allocate workspaces), then the capture.  The returned tensors are
    the graph's static outputs -- valid until the next call with the same shape; `calibrate()` consumes them at once
    (the engine reads them, `_post_process` interpolates them into new tensors).  `graph=False` runs the same code
    eagerly (CPU tests, debugging)."""

    def __init__(self, net: torch.nn.Module, graph: bool = True, warmup: int = 3, capturable_eager: bool = False,
                 max_graphs: int = 8):
        self.net = net
        self.graph = graph
        self.max_graphs = max_graphs  # every graph keeps its activations alive: least recently used shapes are dropped
        self.capturable_eager = capturable_eager  # graph=False only: run the eager modules through the capture-safe path
        self.warmup = warmup
        self._graphs: Dict[Tuple, Tuple[torch.cuda.CUDAGraph, torch.Tensor, Dict[str, torch.Tensor]]] = {}
        self.replays = 0
        self.captures = 0

    def _capture(self, image: torch.Tensor):
        static_in = image.clone()
        side = torch.cuda.Stream(device=image.device)
        side.wait_stream(torch.cuda.current_stream(image.device))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(self.warmup):
                cnn_forward(self.net, static_in, capturable=True)
        torch.cuda.current_stream(image.device).wait_stream(side)
        g = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(g):
            static_out = cnn_forward(self.net, static_in, capturable=True)
        self.captures += 1
        return g, static_in, static_out

    def _lookup(self, key: Tuple, image: torch.Tensor):
        """The graph of this input shape: most recently used last; capture on a miss, drop the oldest beyond the cap."""
        entry = self._graphs.pop(key, None)
        if entry is None:
            while len(self._graphs) >= max(1, self.max_graphs):
                self._graphs.pop(next(iter(self._graphs)))
            entry = self._capture(image.contiguous())
        self._graphs[key] = entry
        return entry

    def __call__(self, image: torch.Tensor) -> Dict[str, torch.Tensor]:
        if not self.graph:
            with torch.no_grad():
                return cnn_forward(self.net, image, capturable=self.capturable_eager)
        if not image.is_cuda:
            raise RuntimeError("GraphedFrontEnd: CUDA tensors only (graph=False runs the modules eagerly)")
        if self.net.training:
            raise RuntimeError("GraphedFrontEnd is an inference path: put the network in eval mode")
        g, static_in, static_out = self._lookup((image.device.index, tuple(image.shape), image.dtype), image)
        static_in.copy_(image)
        g.replay()
        self.replays += 1
        return static_out


def install_frontend(model, graph: bool = True) -> GraphedFrontEnd:
    """Put a `GraphedFrontEnd` in front of the optimizer of a reference extractor (or of its `.model`).

    `net.forward` becomes: graph replay of the CNN, pass-through of the reference's extra keys, `net.optimizer(out)` --
    the statements of geocalib.py:106-119 with the first three module calls replayed instead of launched.  Combine with
    `geocalib_b200.install(model)` so that the optimizer is the engine.  Undo with `uninstall_frontend(model)`."""
    net = getattr(model, "model", model)
    front = GraphedFrontEnd(net, graph=graph)
    undo_bases = use_device_bases(net)

    def forward(data: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
        out = dict(front(data["image"]))
        out |= {k: data[k] for k in PASS_THROUGH if k in data}
        out |= net.optimizer(out)
        return out

    net.forward = forward
    net._gclm_frontend = (front, undo_bases)
    return front


def uninstall_frontend(model) -> None:
    net = getattr(model, "model", model)
    st = net.__dict__.pop("_gclm_frontend", None)
    if st is not None:
        st[1]()
    net.__dict__.pop("forward", None)
