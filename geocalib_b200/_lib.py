"""ctypes binding of libgclm.so (the C ABI declared in include/gclm.h).

The library is built in-tree by `geocalib_b200/csrc/build.sh` (or `__graft_entry__.build()`).
There is deliberately NO fallback: if the shared library is missing or fails to load, importing
the engine raises, so a silent CPU/eager path can never be mistaken for the CUDA engine.
"""

from __future__ import annotations

import ctypes
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_int32, c_int64, c_uint64, c_void_p
from pathlib import Path

CSRC = Path(__file__).resolve().parent / "csrc"
import os as _os

# GCLM_LIB=<path> loads an alternative build of the same ABI (A/B experiments in tools/)
LIB_PATH = Path(_os.environ["GCLM_LIB"]) if _os.environ.get("GCLM_LIB") else CSRC / "libgclm.so"

GCLM_OK = 0
STOP_BATCH, STOP_PER_IMAGE = 0, 1
SCHEDULE_AUTO, SCHEDULE_PER_PASS, SCHEDULE_PERSISTENT = 0, 1, 2
STATUS_CHOLESKY_FAILED, STATUS_COV_SINGULAR, STATUS_NONFINITE, STATUS_COMM_TIMEOUT = 1, 2, 4, 8
COMM_ID_BYTES = 128
IPC_HANDLE_BYTES = 64

CAMERA_MODEL_IDS = {"pinhole": 0, "simple_radial": 1, "radial": 2, "simple_divisional": 3}


class GclmError(RuntimeError):
    """A libgclm call returned a non-zero status."""


class Problem(ctypes.Structure):
    _fields_ = [
        ("batch", c_int32), ("height", c_int32), ("width", c_int32), ("camera_model", c_int32),
        ("up_field", c_void_p), ("latitude_field", c_void_p),
        ("up_confidence", c_void_p), ("latitude_confidence", c_void_p),
        ("camera_init", c_void_p), ("gravity_init", c_void_p),
        ("estimate_gravity", c_int32), ("estimate_focal", c_int32), ("estimate_dist", c_int32),
        ("shared_intrinsics", c_int32),
    ]


class Options(ctypes.Structure):
    _fields_ = [
        ("num_steps", c_int32), ("lambda_", c_float), ("fix_lambda", c_int32),
        ("early_stop", c_int32), ("atol", c_float), ("rtol", c_float),
        ("use_spherical_manifold", c_int32), ("use_log_focal", c_int32),
        ("up_loss_fn_scale", c_float), ("lat_loss_fn_scale", c_float),
        ("stop_mode", c_int32), ("compute_uncertainty", c_int32), ("rows_per_tile", c_int32),
        ("schedule", c_int32), ("reserved", c_int32 * 2),
    ]


class Result(ctypes.Structure):
    _fields_ = [
        ("camera", c_void_p), ("gravity", c_void_p), ("covariance", c_void_p),
        ("uncertainty", c_void_p), ("costs", c_void_p), ("stop_at", c_void_p),
        ("status", c_void_p), ("hessian_rpf", c_void_p),
    ]


class PackedFields(ctypes.Structure):
    _fields_ = [
        ("batch", c_int32), ("height", c_int32), ("width", c_int32), ("conf_format", c_int32),
        ("up", c_void_p), ("sin_lat", c_void_p), ("up_conf", c_void_p), ("lat_conf", c_void_p),
    ]


CONF_F32, CONF_F16 = 0, 1

# every symbol include/gclm.h declares: (restype, argtypes)
SIGNATURES = {
    "gclm_version": (c_int, []),
    "gclm_strerror": (c_char_p, [c_int]),
    "gclm_default_options": (None, [POINTER(Options)]),
    "gclm_num_params": (c_int, [c_int, c_int, c_int]),
    "gclm_create": (c_int, [c_int, POINTER(c_void_p)]),
    "gclm_destroy": (c_int, [c_void_p]),
    "gclm_reserve": (c_int, [c_void_p, c_int, c_int, c_int]),
    "gclm_solve": (c_int, [c_void_p, POINTER(Problem), POINTER(Options), POINTER(Result), c_void_p]),
    "gclm_solve_packed": (c_int, [c_void_p, POINTER(PackedFields), POINTER(Problem), POINTER(Options), POINTER(Result), c_void_p]),
    "gclm_pack_fields": (c_int, [c_void_p, POINTER(Problem), c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gclm_pack_decoder_outputs": (c_int, [c_void_p, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_void_p, c_int,
                                          c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gclm_system": (c_int, [c_void_p, POINTER(Problem), POINTER(Options), c_int, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gclm_adjoint": (c_int, [c_void_p, POINTER(Problem), POINTER(Options), c_void_p, c_void_p, c_void_p,
                             c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "gclm_solve_shared": (c_int, [c_void_p, POINTER(Problem), POINTER(Options), POINTER(Result), c_void_p, c_void_p]),
    "gclm_comm_unique_id": (c_int, [c_void_p]),
    "gclm_comm_create": (c_int, [c_void_p, c_int, c_int, c_int, POINTER(c_void_p)]),
    "gclm_comm_destroy": (c_int, [c_void_p]),
    "gclm_comm_p2p_export": (c_int, [c_void_p, c_void_p]),
    "gclm_comm_p2p_attach": (c_int, [c_void_p, c_void_p]),
    "gclm_comm_p2p_disable": (c_int, [c_void_p]),
    "gclm_solve_host": (c_int, [c_void_p, POINTER(Problem), POINTER(Options), POINTER(Result), c_int]),
    "gclm_launch_count": (c_int64, [c_void_p]),
    "gclm_persistent_count": (c_int64, [c_void_p]),
    "gclm_tile_config": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int, POINTER(c_int32)]),
    "gclm_pass_timing": (c_int, [c_void_p, c_int]),
    "gclm_pass_time": (c_int, [c_void_p, POINTER(ctypes.c_double), POINTER(c_int64)]),
    "gclm_render_fields": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_float,
                                   c_float, c_uint64, c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
}

_lib = None


def build(verbose: bool = False) -> Path:
    """Compile libgclm.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    res = subprocess.run(["bash", str(CSRC / "build.sh")], capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError(f"building libgclm.so failed:\n{res.stdout}\n{res.stderr}")
    if verbose:
        print(res.stdout, res.stderr)
    return LIB_PATH


def load() -> ctypes.CDLL:
    """Load libgclm.so and type every exported entry point.  Raises if it is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise GclmError(
            f"{LIB_PATH} is missing: the CUDA engine is not built (run geocalib_b200/csrc/build.sh "
            "or __graft_entry__.build()).  There is no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    for name, (restype, argtypes) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the ABI and the header drifted apart
        fn.restype = restype
        fn.argtypes = argtypes
    _lib = lib
    return lib


def check(status: int, what: str = "libgclm") -> None:
    if status != GCLM_OK:
        msg = load().gclm_strerror(status).decode()
        raise GclmError(f"{what} failed: {msg} (status {status})")


def default_options() -> Options:
    opt = Options()
    load().gclm_default_options(ctypes.byref(opt))
    return opt
