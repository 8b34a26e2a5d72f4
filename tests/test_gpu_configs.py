"""BASELINE.json configs as parity cases on the GPU (C1 lives in test_gpu_parity.py):
C2 ILMM p=8 m=4 N=2048 at full size against the oracle; C3 OILMM p=64 N=8192 Matern52 (full m=16
for the per-latent terms of three latents, m=4 for the full posterior-marginals comparison);
C4-size single latent (N=16384) against LAPACK; plus size-independent properties at full size."""
import numpy as np
import pytest

from oracle import lmm_oracle as o

pytestmark = pytest.mark.gpu
RTOL = 1e-9


@pytest.fixture(scope="module")
def lmm():
    import lmm_b200

    lmm_b200.default_context()
    return lmm_b200


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


def c3_problem(m, N=8192, p=64, Ns=1024, seed=0):
    rng = np.random.default_rng(seed)
    x = np.sort(rng.uniform(0.0, N / 100.0, N))
    xs = rng.uniform(0.0, N / 100.0, Ns)
    U, S = o.orthogonal_from_seed(p, m, seed=1)
    inv_ls = np.random.default_rng(2).uniform(0.5, 2.0, m)
    fs = [o.GP(o.Kernel(o.MATERN52, 1.0, float(s))) for s in inv_ls]
    y = rng.standard_normal(p * N)
    return x, xs, U, S, fs, y


def to_gp(lmm, g):
    k = {o.SE: lmm.SEKernel, o.MATERN32: lmm.Matern32Kernel, o.MATERN52: lmm.Matern52Kernel}[g.kernel.kind]()
    return lmm.GP(g.mean_const, (g.kernel.variance * k).compose(lmm.ScaleTransform(g.kernel.inv_lengthscale)))


def test_c2_ilmm_full_size(lmm):
    """ILMM p=8, m=4, N=2048: projected (8192 x 8192) form vs oracle; dense (16384 x 16384) form
    agrees with it up to the reference's 1e-9 jitter."""
    rng = np.random.default_rng(0)
    N, p, m = 2048, 8, 4
    x = np.sort(rng.uniform(0, N / 100.0, N))
    H = np.random.default_rng(1).uniform(0, 1, (p, m))
    fs = [o.GP(o.Kernel(k, 1.0, s)) for k, s in zip([o.SE, o.MATERN32, o.MATERN52, o.SE], [0.8, 1.1, 1.4, 1.9])]
    y = rng.standard_normal(p * N)
    f = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in fs]), H)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    got = lmm.logpdf(fx, y)
    assert rel(got, o.ilmm_logpdf(fs, H, x, 0.1, y)) < RTOL
    lmm.set_ilmm_form(1)
    try:
        dense = lmm.logpdf(fx, y)
    finally:
        lmm.set_ilmm_form(0)
    assert rel(dense, got) < 1e-7


def test_c3_oilmm_matern52_terms_full_m(lmm):
    """OILMM p=64, m=16, N=8192 Matern52: per-latent lml terms of latents 0, 7, 15 and the regulariser."""
    m = 16
    x, xs, U, S, fs, y = c3_problem(m)
    f = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    terms = lmm.logpdf_terms(f(lmm.MOInputIsotopicByOutputs(x, 64), 0.1), y)
    T, ST = o.project_orthogonal(U, S, 0.1)
    Y = o.reshape_y(y, len(x))
    for i in (0, 7, 15):
        assert rel(terms[i], o.gp_logpdf(fs[i], x, ST[i], T[i] @ Y)) < RTOL
    assert rel(terms[m], o.regulariser_orthogonal(U, S, 0.1, Y)) < RTOL


def test_c3_oilmm_matern52_posterior_marginals(lmm):
    """Same shape with m=4 latents: logpdf + posterior marginals at N*=1024 vs the oracle."""
    m = 4
    x, xs, U, S, fs, y = c3_problem(m, seed=3)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    post, lp = lmm.posterior(f(lmm.MOInputIsotopicByOutputs(x, 64), 0.1), y, with_logpdf=True)
    assert rel(lp, o.oilmm_logpdf(om, x, 0.1, y)) < RTOL
    M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, 64), 0.1))
    Mr, Vr = o.oilmm_mean_and_var(o.oilmm_posterior(om, x, 0.1, y), xs, 0.1)
    np.testing.assert_allclose(M, Mr, rtol=RTOL, atol=1e-10)
    np.testing.assert_allclose(V, Vr, rtol=RTOL)


def test_c4_size_single_latent_and_properties(lmm):
    """N=16384 SE latent (the C4 factorisation size): lml vs LAPACK, and (K+σ²I)α = δ at full size."""
    rng = np.random.default_rng(0)
    N = 16384
    x = np.sort(rng.uniform(0.0, N / 100.0, N))
    g = o.GP(o.Kernel(o.SE, 1.0, 1.3))
    y = rng.standard_normal(N)
    f = lmm.independent_mogp([to_gp(lmm, g)])
    post, lp = lmm.posterior(f(lmm.MOInputIsotopicByOutputs(x, 1), 0.05), y, with_logpdf=True)
    assert rel(lp, o.gp_logpdf(g, x, 0.05, y)) < RTOL
    alpha = post.fs[0].alpha
    K = o.kernelmatrix(g.kernel, x)
    resid = K @ alpha + 0.05 * alpha - y
    assert np.linalg.norm(resid) / np.linalg.norm(y) < 1e-10
    # posterior mean at the training inputs is y - σ² α (size-independent identity)
    sub = np.arange(0, N, 64)
    M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(x[sub], 1), 0.05))
    np.testing.assert_allclose(M, (y - 0.05 * alpha)[sub], rtol=1e-8, atol=1e-9)
    assert np.all(V > 0.05) and np.all(V < 0.05 + 1.0)
