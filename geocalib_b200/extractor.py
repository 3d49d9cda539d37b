"""Batched `calibrate()` for the reference extractor.

The reference's public entry point, `geocalib.GeoCalib.calibrate` (geocalib/extractor.py:71-127), asserts
B == 1 unless the intrinsics are shared (extractor.py:93-95), and its `_post_process` (extractor.py:51-69) is
written for one image -- so the engine's throughput path (thousands of images per solve) cannot be reached
from the public API.  `calibrate_batch` is that call without the restriction:

  * images are grouped by their original size (one `ImagePreprocessor` geometry per group: same resize,
    `scales`, `crop_pad`; utils.py:88-124), every group goes through the CNN in chunks of `max_batch` and
    through ONE LM solve per chunk;
  * with the B200 optimizer installed (`geocalib_b200.install`) the solve runs with
    `stop_mode="per_image"`, i.e. every image stops exactly where a B == 1 `calibrate()` call would have
    stopped it (the reference's criterion is batch-global, lm_optimizer.py:90-92) -- so the result equals
    looping `calibrate()` over the images;
  * `_post_process` is applied per chunk: `undo_scale_crop`, bilinear resize of the fields back to the input
    size, focal-uncertainty rescaling -- the reference's own code on a batch;
  * results come back in input order, in the reference's container types.

Nothing of the reference is edited; `model` is the reference extractor object (or anything with `.model`,
`.image_processor` and `_post_process`).
"""

from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Union

import torch

from .lm_optimizer import LMOptimizer

TensorOrList = Union[torch.Tensor, Sequence[torch.Tensor]]

FIELD_KEYS = ("up_field", "latitude_field", "up_confidence", "latitude_confidence")


def _as_list(imgs: TensorOrList) -> List[torch.Tensor]:
    if isinstance(imgs, torch.Tensor):
        if imgs.dim() == 3:
            return [imgs]
        assert imgs.dim() == 4, f"expected (B,C,H,W) or (C,H,W), got {tuple(imgs.shape)}"
        return list(imgs.unbind(0))
    return list(imgs)


def _cat_wrapped(items):
    """Concatenate TensorWrapper-like objects (Camera / Gravity: `_data` of shape (b, D))."""
    first = items[0]
    return first.__class__(torch.cat([x._data for x in items], 0))


@torch.no_grad()
def calibrate_batch(
    model,
    imgs: TensorOrList,
    camera_model: str = "pinhole",
    priors: Optional[Dict[str, torch.Tensor]] = None,
    shared_intrinsics: bool = False,
    max_batch: int = 64,
) -> Dict[str, object]:
    """Calibrate many images in one call (reference semantics of `calibrate`, extractor.py:71-127, per image).

    Args:
        model: the reference extractor (`geocalib.GeoCalib`), ideally after `geocalib_b200.install(model)`.
        imgs: (B,3,H,W) tensor or a list of (3,H,W) tensors in [0,1], RGB; sizes may differ.
        camera_model: "pinhole" | "simple_radial" | "radial" | "simple_divisional".
        priors: optional {"focal": (B,) or scalar, "gravity": (B,3) / Gravity}, as in `calibrate`.
        shared_intrinsics: one set of intrinsics for ALL images (then all images must have one size and
            the call is the reference's own shared path, extractor.py:93-95).
        max_batch: images per CNN forward / LM solve (bounds activation memory).

    Returns: the dict `calibrate` returns, batched over the images in input order.  `camera`, `gravity`,
    `covariance` and the `*_uncertainty` entries are concatenated; the field / confidence entries are
    tensors when all images share one size, else lists of per-image tensors.
    """
    images = _as_list(imgs)
    n = len(images)
    assert n > 0
    priors = priors or {}
    prior_focal = priors.get("focal")
    if prior_focal is not None:
        prior_focal = torch.as_tensor(prior_focal)
        prior_focal = prior_focal.expand(n) if prior_focal.dim() == 0 else prior_focal
        assert prior_focal.shape[0] == n
    prior_gravity = priors.get("gravity")
    if prior_gravity is not None:
        g = prior_gravity.vec3d if hasattr(prior_gravity, "vec3d") else torch.as_tensor(prior_gravity)
        prior_gravity = g.expand(n, 3) if g.dim() == 1 else g
        assert prior_gravity.shape[0] == n

    groups: Dict[tuple, List[int]] = {}
    for i, im in enumerate(images):
        assert im.dim() == 3, f"image {i}: expected (C,H,W), got {tuple(im.shape)}"
        groups.setdefault(tuple(im.shape[-2:]), []).append(i)
    if shared_intrinsics:
        assert len(groups) == 1, "shared intrinsics: all images must have the same size"
        max_batch = n  # one estimate for the union of the images

    net = model.model
    optimizer = net.optimizer
    optimizer.set_camera_model(camera_model)
    optimizer.shared_intrinsics = shared_intrinsics
    ours = isinstance(optimizer, LMOptimizer)
    saved_mode = optimizer.conf.stop_mode if ours else None
    if ours and not shared_intrinsics:
        optimizer.conf.stop_mode = "per_image"  # what a loop of B == 1 calls does

    per_image: List[Optional[Dict[str, object]]] = [None] * n
    try:
        for _, idx in groups.items():
            for s0 in range(0, len(idx), max_batch):
                chunk = idx[s0:s0 + max_batch]
                batch = torch.stack([images[i] for i in chunk], 0)
                img_data = model.image_processor(batch)
                prior_values = {}
                if prior_focal is not None:
                    pf = prior_focal[chunk].to(batch)
                    prior_values["prior_focal"] = pf * img_data["scales"][1]
                if prior_gravity is not None:
                    prior_values["prior_gravity"] = prior_gravity[chunk].to(batch)
                out = net(img_data | prior_values)
                camera, out = model._post_process(out["camera"], img_data, out)
                gravity = out["gravity"]
                for j, i in enumerate(chunk):
                    rec = {"camera": camera[j:j + 1], "gravity": gravity[j:j + 1],
                           "covariance": out["covariance"][j:j + 1]}
                    for k, v in out.items():
                        if "field" in k or "confidence" in k or "uncertainty" in k:
                            rec[k] = v[j:j + 1]
                    per_image[i] = rec
    finally:
        if ours:
            optimizer.conf.stop_mode = saved_mode

    result: Dict[str, object] = {
        "camera": _cat_wrapped([r["camera"] for r in per_image]),
        "gravity": _cat_wrapped([r["gravity"] for r in per_image]),
        "covariance": torch.cat([r["covariance"] for r in per_image], 0),
    }
    same_size = len(groups) == 1
    for k in per_image[0]:
        if k in result:
            continue
        vals = [r[k] for r in per_image]
        if "uncertainty" in k or same_size:
            result[k] = torch.cat(vals, 0)
        else:
            result[k] = [v[0] for v in vals]
    return result
