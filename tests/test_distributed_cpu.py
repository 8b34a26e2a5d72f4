"""world_size-2 gloo test (CPU) of the multi-rank host logic: latent shard ranges and the
reduction of per-latent lml terms.  The per-rank compute is stood in for by the CPU oracle (test
infrastructure); on the GPU the same shard ranges are computed inside liblmm and the reduction is
an ncclAllReduce on the compute stream."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import lmm_b200 as lmm
    from oracle import lmm_oracle as o

    rng = np.random.default_rng(0)
    N, p, m = 40, 5, 3
    x = np.sort(rng.uniform(0, 4, N))
    U, S = o.orthogonal_from_seed(p, m, seed=3)
    fs = [o.GP(o.Kernel(k)) for k in (o.SE, o.MATERN32, o.MATERN52)]
    y = rng.standard_normal(N * p)
    model = o.OILMMModel(fs, U, S)
    terms, reg = o.oilmm_logpdf_terms(model, x, 0.1, y)
    lo, hi = lmm.dist.shard_range(m, world, rank)
    local = np.zeros(m + 1)
    local[lo:hi] = terms[lo:hi]  # this rank's latents only
    if rank == 0:
        local[m] = reg  # regulariser computed once (rank 0)
    total = lmm.dist.reduce_terms(local)
    q.put((rank, lo, hi, float(np.sum(total[:m]) + total[m]), float(np.sum(terms) + reg)))
    dist.destroy_process_group()


def test_shard_ranges_cover_all_latents():
    import lmm_b200 as lmm

    for m in (1, 2, 7, 64, 128):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = lmm.dist.shard_range(m, world, r)
                got.extend(range(lo, hi))
            assert got == list(range(m))


def test_gloo_world2_term_reduction():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    res.sort()
    assert (res[0][1], res[0][2]) == (0, 1) and (res[1][1], res[1][2]) == (1, 3)
    for _, _, _, got, ref in res:
        assert got == pytest.approx(ref, rel=1e-14)
