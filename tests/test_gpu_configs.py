"""BASELINE.json configs as parity cases on the GPU (C1 lives in test_gpu_parity.py):
C2 ILMM p=8 m=4 N=2048 at full size against the oracle; C3 OILMM p=64 N=8192 Matern52 (full m=16
for the per-latent terms of three latents, m=4 for the full posterior-marginals comparison);
C4-size single latent (N=16384) against LAPACK; plus size-independent properties at full size."""
import numpy as np
import pytest

from oracle import lmm_oracle as o
from _tol import assert_isapprox, relnorm

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
    assert_isapprox(M, Mr, RTOL)
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
    assert_isapprox(M, (y - 0.05 * alpha)[sub], 1e-9)
    assert np.all(V > 0.05) and np.all(V < 0.05 + 1.0)


# ---- full BASELINE shapes: C4 at p = m = 64, N = 16384; C3 marginals at m = 16; C5-shaped sweep ----
def oracle_latent(g, x, noise, delta, xs):
    """One latent of the oracle with ONE factorisation: lml term, α, posterior mean / variance at xs (the same
    AbstractGPs formulas as o.gp_logpdf / o.gp_posterior / o.gp_mean / o.gp_var, which each factor again)."""
    n = len(x)
    C = o.kernelmatrix(g.kernel, x)
    C[np.diag_indices_from(C)] += noise
    L = o._chol_lower(C)
    del C
    d = np.asarray(delta, dtype=np.float64) - g.mean_const
    z = o._fwd(L, d)
    lml = -0.5 * (n * o.LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))
    alpha = o._bwd(L, z)
    Kxs = o.kernelmatrix(g.kernel, x, xs)
    mean = g.mean_const + Kxs.T @ alpha
    V = o._fwd(L, Kxs)
    var = o.kernelmatrix_diag(g.kernel, xs) - np.sum(V * V, axis=0)
    return lml, alpha, mean, var


def test_c4_full_model_shape(lmm):
    """BASELINE config 4 at its full MODEL shape (p = 64, m = 64, N = 16384, bench.py's synthetic workload): the p x m
    projection and the regulariser, the lml terms of latents 0 / 31 / 63, their α and their posterior mean / variance at
    256 test points, all against the oracle at 1e-9 (means and α norm-relative, no absolute floor)."""
    from bench import workload

    p, m, N, Ns = 64, 64, 16384, 256
    x, U, S, inv_ls, y, s2 = workload(p, m, N)
    xs = np.random.default_rng(7).uniform(0.0, N / 100.0, Ns)
    fs = [o.GP(o.Kernel(o.SE, 1.0, float(s))) for s in inv_ls]
    f = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), s2)
    post, lp = lmm.posterior(fx, y, with_logpdf=True)
    terms = lmm.logpdf_terms(fx, y)
    assert rel(float(np.sum(terms)), lp) < 1e-13
    T, ST = o.project_orthogonal(U, S, s2)
    Y = o.reshape_y(y, N)
    assert rel(terms[m], o.regulariser_orthogonal(U, S, s2, Y)) < RTOL
    Mfull, Vfull = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), s2))
    H = U * np.sqrt(S)[None, :]
    contrib_m, contrib_v = np.zeros((p, Ns)), np.zeros((p, Ns))
    for i in (0, 31, 63):
        lml, alpha, mean, var = oracle_latent(fs[i], x, ST[i], T[i] @ Y, xs)
        assert rel(terms[i], lml) < RTOL, (i, terms[i], lml)
        gi = post.f.fs[i]
        assert_isapprox(gi.delta, T[i] @ Y, 1e-12, f"projection row {i}")
        assert_isapprox(gi.alpha, alpha, RTOL, f"alpha of latent {i}")
        Mi, Vi = lmm.mean_and_var(gi(xs, 0.0))
        assert_isapprox(Mi, mean, RTOL, f"posterior mean of latent {i}")
        np.testing.assert_allclose(Vi, var, rtol=RTOL)
        contrib_m += np.outer(H[:, i], mean)
        contrib_v += np.outer(H[:, i] ** 2, var + 1e-18)
    # the mixed marginals are the H-weighted sums of the latent ones (src/oilmm.jl:69-72): rebuild them from the GPU's own
    # latent marginals for ALL latents and compare with the fused back-projection
    ML = np.stack([lmm.mean_and_var(post.f.fs[i](xs, 0.0))[0] for i in range(m)])
    VL = np.stack([lmm.mean_and_var(post.f.fs[i](xs, 0.0))[1] for i in range(m)])
    assert_isapprox(Mfull, (H @ ML).reshape(-1), 1e-12, "back-projected mean")
    np.testing.assert_allclose(Vfull, ((H * H) @ (VL + 1e-18) + s2).reshape(-1), rtol=1e-12)
    post.f.fs[0]._owner.free()


def test_c3_posterior_marginals_full_m16(lmm):
    """BASELINE config 3 at full size: OILMM p = 64, m = 16, N = 8192 Matern52, logpdf + posterior marginals at N* = 1024
    against the oracle (one CPU factorisation per latent)."""
    m, p, N, Ns = 16, 64, 8192, 1024
    x, xs, U, S, fs, y = c3_problem(m, seed=11)
    f = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    post, lp = lmm.posterior(f(lmm.MOInputIsotopicByOutputs(x, p), 0.1), y, with_logpdf=True)
    M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
    T, ST = o.project_orthogonal(U, S, 0.1)
    Y = o.reshape_y(y, N)
    H = U * np.sqrt(S)[None, :]
    ML, VL, lmls = np.zeros((m, Ns)), np.zeros((m, Ns)), np.zeros(m)
    for i in range(m):
        lmls[i], _, ML[i], VL[i] = oracle_latent(fs[i], x, ST[i], T[i] @ Y, xs)
    assert rel(lp, float(np.sum(lmls)) + o.regulariser_orthogonal(U, S, 0.1, Y)) < RTOL
    assert_isapprox(M, (H @ ML).reshape(-1), RTOL, "C3 posterior mean")
    np.testing.assert_allclose(V, ((H * H) @ (VL + 1e-18) + 0.1).reshape(-1), rtol=RTOL)
    post.f.fs[0]._owner.free()


def test_c5_shaped_sweep(lmm):
    """BASELINE config 5's shape (p = 256, N = 8192, lengthscale sweep): (a) 2 latents x 2 sweep points against the oracle;
    (b) at the full m = 128, each sweep point equals the plain logpdf with the scaled lengthscales, whose lml terms for
    latents 0 and 127 and regulariser are checked against the oracle."""
    p, N = 256, 8192
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(0.0, N / 100.0, N))
    y = rng.standard_normal(p * N)
    scales = np.array([0.5, 2.0])
    O = lmm.MOInputIsotopicByOutputs
    # (a)
    m = 2
    U, S = o.orthogonal_from_seed(p, m, seed=1)
    inv_ls = np.array([0.8, 1.3])
    fs = [o.GP(o.Kernel(o.SE, 1.0, float(s))) for s in inv_ls]
    f = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    got = lmm.logpdf_sweep(f(O(x, p), 0.1), y, scales)
    for k, sc in enumerate(scales):
        om = o.OILMMModel([o.GP(o.Kernel(o.SE, 1.0, float(s * sc))) for s in inv_ls], U, S)
        assert rel(got[k], o.oilmm_logpdf(om, x, 0.1, y)) < RTOL
    # (b)
    m = 128
    U, S = o.orthogonal_from_seed(p, m, seed=1)
    inv_ls = np.random.default_rng(2).uniform(0.5, 2.0, m)
    fs = [o.GP(o.Kernel(o.SE, 1.0, float(s))) for s in inv_ls]
    f = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    got = lmm.logpdf_sweep(f(O(x, p), 0.1), y, scales)
    T, ST = o.project_orthogonal(U, S, 0.1)
    Y = o.reshape_y(y, N)
    reg = o.regulariser_orthogonal(U, S, 0.1, Y)
    for k, sc in enumerate(scales):
        fk = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, o.GP(o.Kernel(o.SE, 1.0, float(s * sc)))) for s in inv_ls]), lmm.Orthogonal(U, S))
        terms = lmm.logpdf_terms(fk(O(x, p), 0.1), y)
        assert rel(got[k], float(np.sum(terms))) < 1e-12
        assert rel(terms[m], reg) < RTOL
        for i in (0, m - 1):
            assert rel(terms[i], o.gp_logpdf(o.GP(o.Kernel(o.SE, 1.0, float(inv_ls[i] * sc))), x, ST[i], T[i] @ Y)) < RTOL
