"""The C-ABI library loads on a CPU-only box and exports every symbol include/gclm.h declares.
No compute calls (there is no GPU here)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from geocalib_b200 import _lib

    if not _lib.LIB_PATH.exists():
        _lib.build()
    return _lib.load()


def declared_symbols():
    text = (ROOT / "include" / "gclm.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gclm_[a-z_0-9]+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from geocalib_b200 import _lib

    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/gclm.h but not exported by libgclm.so"
    assert set(names) == set(_lib.SIGNATURES), set(names) ^ set(_lib.SIGNATURES)


def test_no_gpu_calls_fail_cleanly(lib):
    import torch

    from geocalib_b200 import _lib

    assert lib.gclm_version() == 100
    assert lib.gclm_strerror(0) == b"ok"
    assert lib.gclm_num_params(0, 1, 1) == 3
    assert lib.gclm_num_params(2, 1, 1) == 5
    assert lib.gclm_num_params(3, 0, 1) == 2
    assert lib.gclm_num_params(7, 1, 1) < 0
    opt = _lib.default_options()
    assert (opt.num_steps, opt.early_stop, opt.use_log_focal) == (30, 1, 1)
    assert abs(opt.lambda_ - 0.1) < 1e-7 and abs(opt.atol - 1e-8) < 1e-12
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert lib.gclm_create(0, ctypes.byref(h)) == -6  # GCLM_ERR_NO_DEVICE, no crash


def test_struct_layout_matches_header():
    """sizeof of the ctypes mirrors == what the C compiler computes for include/gclm.h."""
    import subprocess
    import tempfile

    from geocalib_b200 import _lib

    src = '#include <stdio.h>\n#include "gclm.h"\nint main(){printf("%zu %zu %zu\\n", sizeof(gclm_problem), sizeof(gclm_options), sizeof(gclm_result));return 0;}\n'
    with tempfile.TemporaryDirectory() as d:
        c = Path(d) / "s.c"
        c.write_text(src)
        exe = Path(d) / "s"
        subprocess.run(["/usr/bin/gcc", "-I", str(ROOT / "include"), str(c), "-o", str(exe)], check=True)
        sizes = list(map(int, subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()))
    assert sizes == [ctypes.sizeof(_lib.Problem), ctypes.sizeof(_lib.Options), ctypes.sizeof(_lib.Result)]
