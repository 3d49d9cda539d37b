"""geocalib_b200 -- a Blackwell-native (sm_100a) Levenberg-Marquardt calibration engine that drops
in behind cvg/GeoCalib's `LMOptimizer` / `Camera` / `Gravity` API.

    from geocalib_b200 import LMOptimizer, install
    install(reference_geocalib_model)            # attribute swap of model.optimizer
    out = LMOptimizer({"camera_model": "pinhole"}).eval()(fields_on_cuda)
    res = calibrate_batch(reference_geocalib_model, images)   # calibrate() without the B == 1 restriction

The hot path lives in `csrc/` (CUDA + C ABI, built to csrc/libgclm.so); importing the package does
not need a GPU, running the optimizer does.  There is no CPU fallback.
"""

from .autograd import differentiable_calibrate
from .camera import BaseCamera, Pinhole, Radial, SimpleDivisional, SimpleRadial, camera_models
from .extractor import calibrate_batch
from .frontend import GraphedFrontEnd, install_frontend, uninstall_frontend
from .gravity import Gravity
from .lm_optimizer import B200LMOptimizer, LMOptimizer, get_trivial_estimation, install

__all__ = [
    "BaseCamera", "Pinhole", "Radial", "SimpleDivisional", "SimpleRadial", "camera_models",
    "Gravity", "LMOptimizer", "B200LMOptimizer", "get_trivial_estimation", "install", "calibrate_batch", "differentiable_calibrate",
    "GraphedFrontEnd", "install_frontend", "uninstall_frontend",
]
__version__ = "0.1.0"
