"""GPU parity of the optional integer-slice (Ozaki) trailing update (csrc/ozaki.cu: int8 tcgen05 MMAs with exact int32 accumulation in
TMEM): the factor against LAPACK and the DMMA path, and an OILMM logpdf + posterior + marginals against the CPU oracle at the north
star's 1e-9.  DMMA stays the default; these tests switch the option on and off again."""
import os

import numpy as np
import pytest
import scipy.linalg as sla

from oracle import lmm_oracle as o
from _tol import assert_isapprox, relnorm
from test_gpu_parity import make_problem, to_lmm_gp

pytestmark = pytest.mark.gpu


@pytest.fixture()
def lmm():
    import lmm_b200

    ctx = lmm_b200.default_context()
    yield lmm_b200
    ctx.set_option("ozaki", int(os.environ.get("LMM_OZAKI", "0")))  # the library default, or what the whole run was started with
    ctx.set_option("ozaki_bits", int(os.environ.get("LMM_OZAKI_BITS", "7")))
    ctx.set_option("ozaki_min_k", 4)
    ctx.set_option("outer_block", 0)


def spd(N, batch, seed, cond_boost=0.0):
    rng = np.random.default_rng(seed)
    A = np.empty((batch, N, N))
    for b in range(batch):
        M = rng.standard_normal((N, N))
        A[b] = M @ M.T / N + (1.0 + cond_boost) * np.eye(N)
        d = np.exp(rng.uniform(-3, 3, N))  # rows of very different scale: one exponent per row has to cope
        A[b] = A[b] * d[:, None] * d[None, :]
    return A


@pytest.mark.parametrize("S,bits,tol", [(8, 7, 5e-14), (7, 7, 2e-11), (6, 7, 2e-9), (7, 8, 5e-13), (6, 8, 1e-10)])
def test_factor_matches_lapack(lmm, S, bits, tol):
    """Normwise relative error of L per matrix (rows scaled over e^-3 .. e^3); S = 8 truncates at 2^-56 (FP64 level: measured 2.1e-14
    against 1.6e-14 for DMMA), every plane less costs 2^7 (measured 2.5e-12 and 3.1e-10).  Radix 256 (bits = 8, balanced digits in
    [-128, 127]): 7 planes carry 54 bits with 28 instead of 36 MMAs per tile product (measured 1.0e-13; 6 planes: 2.5e-11)."""
    ctx = lmm.default_context()
    N, batch = 2600, 3  # 21 tile rows: wide updates at s0 = 8 and 16 take the int8 path (K = 8 and 16 k-tiles)
    A = spd(N, batch, seed=S)
    ctx.set_option("ozaki", 0)
    L0, ld0, info0 = lmm.potrf_batched(A)
    ctx.set_option("ozaki", S)
    ctx.set_option("ozaki_bits", bits)
    L1, ld1, info1 = lmm.potrf_batched(A)
    assert not info0.any() and not info1.any()
    for b in range(batch):
        Lr = sla.cholesky(A[b], lower=True)
        e0, e1 = relnorm(L0[b], Lr), relnorm(L1[b], Lr)
        assert e0 < 2e-14, e0
        assert e1 < tol, (S, bits, e1)
        assert abs(ld1[b] - ld0[b]) <= max(1e-11, tol) * abs(ld0[b]) + 1e-9


def test_small_k_and_block_widths(lmm):
    """Odd / even tile-row counts, block widths 2 .. 8, wide updates over as few as one k-tile."""
    ctx = lmm.default_context()
    for N in (1500, 1700):
        A = spd(N, 3, seed=3)  # batch >= 3: smaller batches take the latency-optimised right-looking DMMA schedule
        Lr = [sla.cholesky(a, lower=True) for a in A]
        for ob, mink in ((2, 1), (3, 2), (5, 4), (0, 4), (8, 8)):
            ctx.set_option("outer_block", ob)
            ctx.set_option("ozaki_min_k", mink)
            for S, bits in ((8, 7), (7, 8)):
                ctx.set_option("ozaki", S)
                ctx.set_option("ozaki_bits", bits)
                L, ld, info = lmm.potrf_batched(A)
                assert not info.any()
                for b in range(3):
                    assert relnorm(L[b], Lr[b]) < 5e-13, (N, ob, mink, S, bits)


def test_not_positive_definite_is_reported(lmm):
    ctx = lmm.default_context()
    A = spd(2000, 3, seed=5)
    A[1, 1500, 1500] = -1.0
    ctx.set_option("ozaki", 8)
    L, ld, info = lmm.potrf_batched(A)
    assert info[0] == 0 and info[1] == 1501 and info[2] == 0


def test_oilmm_against_oracle(lmm):
    """logpdf, posterior marginals: the north star's 1e-9 with the int8 trailing update (N = 3000: 24 tile rows)."""
    ctx = lmm.default_context()
    N, p, m, Ns = 3000, 6, 5, 300  # 24 tile rows of the factor, 3 (ragged) tile rows of test points: the int8 path also runs the prediction sweep
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=21)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    O = lmm.MOInputIsotopicByOutputs
    om = o.OILMMModel(fs, U, S)
    lp_ref = o.oilmm_logpdf(om, x, 0.05, y)
    Mr, Vr = o.oilmm_mean_and_var(o.oilmm_posterior(om, x, 0.05, y), xs, 0.05)
    for Sn, bits in ((8, 7), (7, 7), (7, 8)):
        ctx.set_option("ozaki", Sn)
        ctx.set_option("ozaki_bits", bits)
        post, lp = lmm.posterior(f(O(x, p), 0.05), y, with_logpdf=True)
        M, V = lmm.mean_and_var(post(O(xs, p), 0.05))
        assert abs(lp - lp_ref) <= 1e-9 * abs(lp_ref), (Sn, bits, lp, lp_ref)
        assert_isapprox(M, Mr, 1e-9, f"posterior mean, {Sn} planes of {bits} bits")
        assert_isapprox(V, Vr, 1e-9, f"posterior variance, {Sn} planes of {bits} bits")
