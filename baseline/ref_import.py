"""Import the UNMODIFIED reference package installed under baseline/_ref (baseline/install_ref.sh).

`geocalib/utils.py:9` imports kornia at module top; kornia is not in this image and is only used by
`ImagePreprocessor.resize` (utils.py:128-136).  An empty stand-in module makes the import succeed;
`resize_stub()` gives that one function an antialiased bilinear resize from torch for the tests
that run `calibrate()` end to end.  Nothing here touches the LM path that is being measured.
"""

from __future__ import annotations

import sys
import types
from pathlib import Path

REF_DIR = Path(__file__).resolve().parent / "_ref"


def available() -> bool:
    return (REF_DIR / "geocalib" / "lm_optimizer.py").exists()


def _kornia_stub():
    if "kornia" in sys.modules:
        return
    try:
        import kornia  # noqa: F401
        return
    except Exception:
        pass
    import torch.nn.functional as F

    k = types.ModuleType("kornia")
    geometry = types.ModuleType("kornia.geometry")
    transform = types.ModuleType("kornia.geometry.transform")

    def resize(img, size, side="short", antialias=True, align_corners=None, interpolation="bilinear"):
        h, w = img.shape[-2:]
        if isinstance(size, int):
            if side == "short":
                s = size / min(h, w)
            elif side == "long":
                s = size / max(h, w)
            elif side == "vert":
                s = size / h
            else:
                s = size / w
            size = (int(round(h * s)), int(round(w * s)))
        x = img if img.dim() == 4 else img[None]
        y = F.interpolate(x, size=tuple(size), mode=interpolation, antialias=antialias,
                          align_corners=align_corners if interpolation != "nearest" else None)
        return y if img.dim() == 4 else y[0]

    transform.resize = resize
    geometry.transform = transform
    k.geometry = geometry
    sys.modules["kornia"] = k
    sys.modules["kornia.geometry"] = geometry
    sys.modules["kornia.geometry.transform"] = transform


def import_reference():
    """Returns the reference `geocalib` package (from baseline/_ref); raises if it is not installed."""
    if not available():
        raise ImportError(f"{REF_DIR} is empty: run baseline/install_ref.sh in the build container")
    _kornia_stub()
    if str(REF_DIR) not in sys.path:
        sys.path.insert(0, str(REF_DIR))
    import geocalib  # noqa: F401

    assert Path(geocalib.__file__).resolve().is_relative_to(REF_DIR), geocalib.__file__
    return geocalib
