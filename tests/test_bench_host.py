"""bench.py's host side, on CPU: the reference arm's JSON contract (the driver computes the speed-up from this line and
ours, so `metric`, `unit` and `config` must be identical), and the CPU-count logic behind its process / thread split."""
import json
import os
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "baseline"))
import ref_import  # noqa: E402


def test_effective_cpus_respects_affinity_and_cgroup_quota():
    import bench

    n = bench.effective_cpus()
    assert isinstance(n, int) and 1 <= n <= (os.cpu_count() or 1)
    assert n <= len(os.sched_getaffinity(0))


@pytest.mark.skipif(not ref_import.available(), reason="baseline/_ref is not installed")
def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference --steps 1 --warmup 1`: the unmodified reference optimizer on this host's CPUs."""
    p = subprocess.run([sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"],
                       capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert p.returncode == 0, p.stderr[-2000:]
    line = json.loads(p.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and "unavailable" not in line
    assert line["metric"] == "images/sec (batched LM calibrate)" and line["unit"] == "images/sec"
    assert line["higher_is_better"] is True and line["n_gpus"] == 1 and line["steps"] == 1 and line["warmup"] == 1
    assert line["value"] > 0 and line["ms_per_step"] > 0
    cb = line["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["cores"] >= 1 and cb["value"] == line["value"] and "LMOptimizer" in cb["sample"]
    assert line["e2e"] == {"value": line["value"], "unit": "images/sec", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    # the same `config` dict as the line our own arm printed on a B200 (profiles/): the driver divides the two values
    ours = json.loads((ROOT / "profiles" / "r2_bench_default_final.json").read_text().strip().splitlines()[-1])
    assert line["config"] == ours["config"], (line["config"], ours["config"])
    assert line["dtype"] == ours["dtype"] and line["data"] == ours["data"] and line["scaling"] == ours["scaling"]
