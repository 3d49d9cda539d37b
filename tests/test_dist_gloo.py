"""N > 1 host logic on CPU: world_size-2 `gloo` processes.

Covers (i) the contiguous image sharding used by every multi-GPU mode, (ii) the byte broadcast
that distributes the NCCL unique id, and (iii) the Schur-complement reduction of shared
intrinsics: each rank reduces ITS images to the <= 2 ni^2 + 2 ni partial sums, one all-reduce,
every rank solves the ni x ni system -- and the result equals the reference's dense
(2B + ni)^2 arrow system (lm_optimizer.py:350-383) solved in one piece.
"""
import os
import socket
import sys
from pathlib import Path

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, case, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from geocalib_b200 import parallel
        from oracle import oracle

        g = dict(np.load(ROOT / "tests" / "golden" / f"lm_{case}.npz"))
        model = str(g["model"])
        B, _, h, w = g["in_up_field"].shape
        lam = 0.1
        # (ii) byte broadcast (what make_comm uses for the NCCL id)
        payload = bytes(range(128)) if rank == 0 else b""
        got = parallel.broadcast_bytes(payload, 128, 0)
        assert got == bytes(range(128))
        # (i) sharding
        lo, hi = parallel.shard_range(B, world, rank)
        cam0, g0 = oracle.trivial_init(B, h, w)
        ni = 1 + oracle.NUM_DIST[model]
        part = np.zeros(2 * ni * ni + 2 * ni)
        systems = {}
        for b in range(lo, hi):
            G, H, _ = oracle.system(model, cam0[b], g0[b], g["in_up_field"][b], g["in_latitude_field"][b],
                                    g["in_up_confidence"][b], g["in_latitude_confidence"][b], dtype=np.float64)
            systems[b] = (H, G)
            part += parallel.schur_partials(H, G, lam)
        # (iii) ONE all-reduce of the partial sums
        t = torch.from_numpy(part)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        dc = parallel.schur_solve(t.numpy(), ni, lam)
        dg = {b: parallel.schur_backsub(H, G, dc, lam) for b, (H, G) in systems.items()}
        gathered = [None] * world
        dist.all_gather_object(gathered, (dc, dg, systems))
        if rank == 0:
            q.put(gathered)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("case", ["shared_pinhole", "shared_radial"])
def test_schur_allreduce_equals_dense_arrow_system(case):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, case, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:  # join first: a crashed worker must fail the test, not hang it on q.get()
        p.join(120)
        if p.is_alive():
            p.terminate()
        assert p.exitcode == 0, f"gloo worker exited with {p.exitcode}"
    assert not q.empty()
    gathered = q.get()
    # both ranks solved the same shared step
    np.testing.assert_array_equal(gathered[0][0], gathered[1][0])
    dc = gathered[0][0]
    dg, systems = {}, {}
    for _, d, s in gathered:
        dg.update(d)
        systems.update(s)
    B, ni, lam = len(systems), len(dc), 0.1
    # dense reference system, exactly as lm_optimizer.py:370-383 lays it out
    M = 2 * B + ni
    A, rhs = np.zeros((M, M)), np.zeros(M)
    for b in range(B):
        H, G = systems[b]
        A[2 * b:2 * b + 2, 2 * b:2 * b + 2] = H[:2, :2]
        A[2 * b:2 * b + 2, 2 * B:] = H[:2, 2:]
        A[2 * B:, 2 * b:2 * b + 2] = H[2:, :2]
        A[2 * B:, 2 * B:] += H[2:, 2:]
        rhs[2 * b:2 * b + 2] = G[:2]
        rhs[2 * B:] += G[2:]
    A[np.diag_indices(M)] += np.maximum(np.diag(A) * lam, 1e-6)
    delta = np.linalg.solve(A, rhs)
    np.testing.assert_allclose(dc, delta[2 * B:], rtol=1e-9, atol=1e-12)
    for b in range(B):
        np.testing.assert_allclose(dg[b], delta[2 * b:2 * b + 2], rtol=1e-8, atol=1e-12)


def test_shard_range_partitions_the_batch():
    from geocalib_b200.parallel import shard_range

    for B in (1, 7, 8, 1024, 16384, 16385):
        for world in (1, 2, 3, 8):
            spans = [shard_range(B, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)
