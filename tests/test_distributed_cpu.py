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


# ---- distributed-storage row-cyclic factorisation (csrc/host_chol.cu: chol_factor_rowcyclic_dist): index map and schedule
def _rowcyclic_problem(n, seed=4):
    rng = np.random.default_rng(seed)
    M = rng.standard_normal((n, n))
    return M @ M.T / n + np.eye(n), rng.standard_normal(n)


def test_cyclic_index_is_a_packing():
    from tools.rowcyclic_dist_model import cyc_tile_index, cyc_tiles

    for G in (1, 2, 3, 8):
        for nrows in (1, 5, 17, 40):
            total = 0
            for r in range(G):
                idx = [cyc_tile_index(I, J, G, r) for I in range(r, nrows, G) for J in range(I + 1)]
                assert idx == list(range(len(idx)))  # row after row, no gap, no overlap
                assert len(idx) == cyc_tiles(nrows, G, r)
                total += len(idx)
            assert total == nrows * (nrows + 1) // 2


@pytest.mark.parametrize("G,ob,nt", [(2, 2, 9), (3, 4, 10), (4, 1, 9), (2, 3, 8), (8, 2, 20), (8, 4, 19), (8, 3, 17)])
def test_rowcyclic_dist_schedule_single_process(G, ob, nt):
    """All G ranks simulated in one process: factor, logdet and z = L^{-1} rhs against NumPy; no rank ever holds more than its rows."""
    import scipy.linalg as sla

    from tools.rowcyclic_dist_model import RankState, cyc_tiles, factor

    T = 3
    A, rhs = _rowcyclic_problem(nt * T)
    states = [RankState(A, rhs, T, G, r) for r in range(G)]
    factor(states, ob, lambda bufs: [bufs for _ in bufs])
    L = np.linalg.cholesky(A)
    z = sla.solve_triangular(L, rhs, lower=True)
    for st in states:
        assert st.store.shape[0] == cyc_tiles(nt + 1, G, st.r)
        np.testing.assert_allclose(st.logdet, 2 * np.sum(np.log(np.diag(L))), rtol=1e-12)
        np.testing.assert_allclose(st.z, z, rtol=1e-9, atol=1e-12)
        for I in range(st.r, nt, G):
            for J in range(I + 1):
                ref = L[I * T:(I + 1) * T, J * T:(J + 1) * T]
                np.testing.assert_allclose(np.tril(st.tile(I, J)) if I == J else st.tile(I, J), ref, rtol=1e-9, atol=1e-12)


def _rowcyclic_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from tools.rowcyclic_dist_model import RankState, factor

    T, nt, ob = 3, 9, 2
    A, rhs = _rowcyclic_problem(nt * T)
    st = RankState(A, rhs, T, world, rank)

    def exchange(bufs):  # the ncclAllGather of the CUDA path, here over gloo
        mine = torch.from_numpy(np.nan_to_num(bufs[0], nan=0.0))
        allb = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine)
        return [[b.numpy() for b in allb]]

    factor([st], ob, exchange)
    q.put((rank, float(st.logdet), st.z.copy()))
    dist.destroy_process_group()


def test_gloo_world2_rowcyclic_distributed_storage():
    import scipy.linalg as sla

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_rowcyclic_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=120) for _ in procs]
    for pr in procs:
        pr.join(timeout=60)
        assert pr.exitcode == 0
    A, rhs = _rowcyclic_problem(27)
    L = np.linalg.cholesky(A)
    z = sla.solve_triangular(L, rhs, lower=True)
    for _, logdet, zz in res:  # every rank ends with the full logdet and the full z, from 1/2 of the matrix each
        assert logdet == pytest.approx(2 * np.sum(np.log(np.diag(L))), rel=1e-12)
        np.testing.assert_allclose(zz, z, rtol=1e-9, atol=1e-12)
