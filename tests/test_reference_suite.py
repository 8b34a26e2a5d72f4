"""The reference's own test files restated against the B200 path, testset by testset:
test/oilmm.jl, test/ilmm.jl, test/independent_mogp.jl, test/orthogonal_matrix.jl, with the helpers of
test/test_utils.jl and runtests.jl (`generate_toy_data`, `test_sampling_consistency`, `approx_equivalent`,
`_is_approx`) and a restatement of the AbstractGPs.TestUtils public-interface checks the reference runs
(`test_finitegp_primary_and_secondary_public_interface`, `test_internal_abstractgps_interface`).

Where the reference compares against an independent Julia object (a dense `GP(LinearMixingModelKernel(...))`, single-output
GPs) the CPU oracle's dense / single-GP functions play that role.  Tolerance: Julia's `isapprox` default
(rtol = sqrt(eps)) where the reference uses `≈`."""
import math

import numpy as np
import pytest

from oracle import lmm_oracle as o

pytestmark = pytest.mark.gpu

RTOL = math.sqrt(np.finfo(np.float64).eps)  # Julia isapprox default


@pytest.fixture(scope="module")
def lmm():
    import lmm_b200

    lmm_b200.default_context()
    return lmm_b200


KMAP = {o.SE: "SEKernel", o.MATERN32: "Matern32Kernel", o.MATERN52: "Matern52Kernel"}


def to_gp(lmm, g: o.GP):
    k = getattr(lmm, KMAP[g.kernel.kind])()
    if g.kernel.variance != 1.0:
        k = g.kernel.variance * k
    if g.kernel.inv_lengthscale != 1.0:
        k = k.compose(lmm.ScaleTransform(g.kernel.inv_lengthscale))
    return lmm.GP(g.mean_const, k) if g.mean_const != 0.0 else lmm.GP(k)


def approx(a, b, rtol=RTOL, atol=0.0):
    """Julia `isapprox(a, b)`: norm(a - b) <= max(atol, rtol * max(norm(a), norm(b)))."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.linalg.norm(a - b) <= max(atol, rtol * max(np.linalg.norm(a), np.linalg.norm(b)))


def is_approx_marginals(ma, mb):  # runtests.jl `_is_approx`
    return approx([d.mu for d in ma], [d.mu for d in mb]) and approx([d.sigma for d in ma], [d.sigma for d in mb])


def generate_toy_data(rng):
    """test/test_utils.jl:1-35: 5 points on [0, 10], three outputs drawn from GP(SEKernel())(x, 1e-6), 3/2 split."""
    x = np.linspace(0.0, 10.0, 5)
    K = o.kernelmatrix(o.Kernel(o.SE), x) + 1e-6 * np.eye(5)
    ys = np.linalg.cholesky(K) @ rng.standard_normal((5, 3))
    idx = rng.permutation(5)
    tr, te = idx[:3], idx[3:]
    return x[tr], x[te], ys[tr].T.reshape(-1), ys[te].T.reshape(-1)


def check_sampling_consistency(lmm, rng, f, x_mo, rtol=1e-2, atol=1e2, s2=1e-6):
    """test/test_utils.jl:41-48."""
    fx = f(x_mo, s2)
    y = lmm.rand(rng, fx)
    post = lmm.posterior(fx, y)
    assert approx(lmm.rand(rng, post(x_mo, s2)), y, rtol=rtol)
    assert approx(lmm.mean(post(x_mo)), y, rtol=rtol)
    assert approx(lmm.var(post(x_mo)), np.zeros(len(y)), rtol=rtol, atol=atol)


def check_public_interface(lmm, rng, fx, secondary=True, jitter=1e-10):
    """AbstractGPs.TestUtils.test_finitegp_primary[_and_secondary]_public_interface, restated: shapes, types and the
    self-consistency of rand / marginals / mean / var / cov / mean_and_cov / logpdf / posterior."""
    n = len(fx)
    y = lmm.rand(rng, fx)
    assert y.shape == (n,)
    Y = lmm.rand(rng, fx, 3)
    assert Y.shape == (n, 3)
    ms = lmm.marginals(fx)
    assert len(ms) == n
    M, V = lmm.mean_and_var(fx)
    assert approx([d.mu for d in ms], M) and approx([d.sigma ** 2 for d in ms], V)
    M2, Cm = lmm.mean_and_cov(fx)
    assert Cm.shape == (n, n) and approx(M2, M)
    assert approx(Cm, Cm.T) and approx(np.diag(Cm), V)
    assert np.min(np.linalg.eigvalsh((Cm + Cm.T) / 2)) > -jitter
    lp = lmm.logpdf(fx, y)
    assert isinstance(lp, float) and np.isfinite(lp)
    lps = lmm.logpdf(fx, Y)
    assert lps.shape == (3,) and approx(lps[0], lmm.logpdf(fx, Y[:, 0]))
    assert isinstance(lmm.posterior(fx, y), lmm.api.AbstractGP)
    if secondary:
        assert approx(lmm.mean(fx), M) and approx(lmm.var(fx), V) and approx(lmm.cov(fx), Cm)


def run_test_oilmm(lmm, rng, kernels, U, S, x_train, x_test, y_train, y_test):
    """test/oilmm.jl:1-38 `test_oilmm`."""
    p = U.shape[0]
    fs = lmm.independent_mogp([lmm.GP(k) for k in kernels])
    H = lmm.Orthogonal(U, S)
    ilmm = lmm.ILMM(fs, np.asarray(H))  # collect(H)
    oilmm = lmm.ILMM(fs, H)
    assert isinstance(oilmm, lmm.OILMM)
    O = lmm.MOInputIsotopicByOutputs
    ilmmx, oilmmx = ilmm(O(x_train, p), 0.1), oilmm(O(x_train, p), 0.1)
    assert approx(lmm.mean(ilmmx), lmm.mean(oilmmx))
    assert approx(lmm.var(ilmmx), lmm.var(oilmmx))
    assert approx(lmm.cov(ilmmx), lmm.cov(oilmmx))
    assert approx(lmm.logpdf(ilmmx, y_train), lmm.logpdf(oilmmx, y_train))
    assert is_approx_marginals(lmm.marginals(ilmmx), lmm.marginals(oilmmx))
    assert len(lmm.rand(rng, oilmmx)) == p * len(x_train)
    p_ilmmx, p_oilmmx = lmm.posterior(ilmmx, y_train), lmm.posterior(oilmmx, y_train)
    pi, po = p_ilmmx(O(x_test, p), 0.1), p_oilmmx(O(x_test, p), 0.1)
    assert approx(lmm.mean(pi), lmm.mean(po))
    assert approx(lmm.var(pi), lmm.var(po))
    assert approx(lmm.logpdf(pi, y_test), lmm.logpdf(po, y_test))
    assert is_approx_marginals(lmm.marginals(pi), lmm.marginals(po))
    assert len(lmm.rand(rng, po)) == p * len(x_test)
    check_sampling_consistency(lmm, rng, oilmm, O(x_train, p))
    assert isinstance(lmm.logpdf_and_gradient(oilmmx, y_train, with_grad_y=True), tuple)  # gradient(logpdf, oilmmx, y_train) isa Tuple
    assert isinstance(lmm.logpdf_and_gradient(po, y_test, with_grad_y=True), tuple)       # gradient(logpdf, po, y_test) isa Tuple
    check_public_interface(lmm, rng, oilmmx)
    check_public_interface(lmm, rng, po)


def svd_H(rng, p, m):
    U, S, _ = np.linalg.svd(rng.uniform(0, 1, (p, m)), full_matrices=False)  # test/oilmm.jl:45-46
    return np.ascontiguousarray(U), np.ascontiguousarray(S)


@pytest.mark.parametrize("m,kinds", [(3, ["SE", "M32", "M32"]), (2, ["SE", "M32"]), (1, ["SE"])],
                         ids=["Full Rank, Dense H", "M Latent Processes", "1 Latent Processes"])
def test_oilmm_testsets(lmm, m, kinds):
    """test/oilmm.jl:40-66."""
    rng = np.random.default_rng(4161999)
    x_train, x_test, y_train, y_test = generate_toy_data(rng)
    U, S = svd_H(rng, 3, m)
    kernels = [lmm.SEKernel() if k == "SE" else lmm.Matern32Kernel() for k in kinds]
    run_test_oilmm(lmm, rng, kernels, U, S, x_train, x_test, y_train, y_test)


def run_test_ilmm(lmm, rng, okinds, H, x_train, x_test, y_train, y_test):
    """test/ilmm.jl:1-40 `test_ilmm`: the ILMM against `GP(LinearMixingModelKernel(kernels, H'))` -- here the oracle's dense
    (pN x pN) multi-output GP."""
    p = H.shape[0]
    ofs = [o.GP(o.Kernel(k)) for k in okinds]
    ilmm = lmm.ILMM(lmm.independent_mogp([to_gp(lmm, g) for g in ofs]), H)
    O = lmm.MOInputIsotopicByOutputs
    s2 = 1e-6
    ilmmx = ilmm(O(x_train, p), s2)
    n = p * len(x_train)
    Cn = o.dense_mogp_cov(ofs, H, x_train) + s2 * np.eye(n)
    assert approx(lmm.mean(ilmmx), o.dense_mogp_mean(ofs, H, x_train), atol=1e-12)
    assert approx(lmm.var(ilmmx), np.diag(Cn))
    assert approx(lmm.cov(ilmmx), Cn)
    assert approx(lmm.logpdf(ilmmx, y_train), o.dense_mogp_logpdf(ofs, H, x_train, s2, y_train))
    assert len(lmm.rand(rng, ilmmx)) == n
    p_ilmmx = lmm.posterior(ilmmx, y_train)
    pi = p_ilmmx(O(x_test, p), s2)
    Mn, Vn = o.dense_mogp_posterior_mean_and_var(ofs, H, x_train, s2, y_train, x_test, s2)
    assert approx(lmm.mean(pi), Mn)
    assert approx(lmm.var(pi), Vn, rtol=1e-6)  # the projected form carries src/ilmm.jl:63's 1e-9 jitter
    ms = lmm.marginals(pi)
    assert approx([d.mu for d in ms], Mn) and approx([d.sigma for d in ms], np.sqrt(Vn), rtol=1e-6)
    assert len(lmm.rand(rng, pi)) == p * len(x_test)
    assert np.isfinite(lmm.logpdf(pi, y_test))
    check_sampling_consistency(lmm, rng, ilmm, O(x_train, p))
    assert isinstance(lmm.logpdf_and_gradient(ilmmx, y_train, with_grad_y=True), tuple)
    assert isinstance(lmm.logpdf_and_gradient(pi, y_test, with_grad_y=True), tuple)
    check_public_interface(lmm, rng, ilmmx)
    check_public_interface(lmm, rng, pi, jitter=1e-8)


@pytest.mark.parametrize("m,kinds", [(3, [o.SE, o.MATERN32, o.MATERN32]), (2, [o.SE, o.MATERN32]), (1, [o.SE])],
                         ids=["Full Rank, Dense H", "M Latent Processes", "1 Latent Processes"])
def test_ilmm_testsets(lmm, m, kinds):
    """test/ilmm.jl:42-80."""
    rng = np.random.default_rng(4161999)
    x_train, x_test, y_train, y_test = generate_toy_data(rng)
    H = rng.uniform(0, 1, (3, m))
    run_test_ilmm(lmm, rng, kinds, H, x_train, x_test, y_train, y_test)


def test_ilmm_util(lmm):
    """test/ilmm.jl:55-70 "util": noise_var, reshape_y, unpack, get_latent_gp."""
    rng = np.random.default_rng(1)
    fs = lmm.independent_mogp([lmm.GP(lmm.Matern32Kernel())])
    H = rng.uniform(0, 1, (2, 1))
    x = lmm.MOInputIsotopicByOutputs(lmm.ColVecs(rng.uniform(0, 1, (2, 2))), 2)
    ilmm = lmm.ILMM(fs, H)
    ilmmx = ilmm(x, 0.1)
    assert lmm.noise_var(ilmmx) == 0.1
    y = rng.uniform(0, 1, 16)
    assert lmm.reshape_y(y, 8).shape == (2, 8) and lmm.reshape_y(y, 2).shape == (8, 2)
    lat, Hu, s2, xx = lmm.unpack(ilmmx)
    assert lat is fs and np.array_equal(Hu, H) and s2 == 0.1 and xx is x.x
    assert lmm.get_latent_gp(ilmm) is fs
    with pytest.raises(RuntimeError, match="out dim of x != out dim of f."):  # src/ilmm.jl:52
        lmm.unpack(ilmm(lmm.MOInputIsotopicByOutputs(np.zeros(3), 5), 0.1))


def test_orthogonal_matrix(lmm):
    """test/orthogonal_matrix.jl."""
    rng = np.random.default_rng(0)
    U, S, _ = np.linalg.svd(rng.uniform(0, 1, (4, 3)), full_matrices=False)
    H = lmm.Orthogonal(U, S)
    assert H.shape == (4, 3)
    assert approx(np.asarray(H), U @ np.diag(np.sqrt(S)))
    with pytest.raises(ValueError, match="`U` is not an orthogonal matrix"):  # ArgumentError, src/orthogonal_matrix.jl:22
        lmm.Orthogonal(rng.uniform(0, 1, (4, 3)), S)
    lmm.Orthogonal(rng.uniform(0, 1, (4, 3)), S, validate_fields=False)  # validation can be switched off


def test_independent_mogp_by_outputs(lmm):
    """test/independent_mogp.jl:2-78 "MOInputIsotopicByOutputs": against the two single-output GPs."""
    rng = np.random.default_rng(123)
    x = np.linspace(1, 2, 5)
    eps = 0.5 * rng.standard_normal(5)
    y_1 = 30 + np.sqrt(x) * np.sin(x) + eps
    y_2 = 10 + np.cbrt(x) * np.cos(2 * x) + eps
    idx = rng.permutation(5)
    tr, te = idx[:3], idx[3:]
    x_train, x_test = x[tr], x[te]
    y_train, y_test = np.concatenate([y_1[tr], y_2[tr]]), np.concatenate([y_1[te], y_2[te]])
    g1, g2 = o.GP(o.Kernel(o.MATERN32), 30.0), o.GP(o.Kernel(o.SE), 10.0)
    f = lmm.independent_mogp([to_gp(lmm, g1), to_gp(lmm, g2)])
    O = lmm.MOInputIsotopicByOutputs
    fx = f(O(x_train, 2), 0.1)
    assert approx(lmm.logpdf(fx, y_train), o.gp_logpdf(g1, x_train, 0.1, y_1[tr]) + o.gp_logpdf(g2, x_train, 0.1, y_2[tr]))
    assert approx(lmm.mean(fx), np.concatenate([np.full(3, 30.0), np.full(3, 10.0)]))
    assert approx(lmm.var(fx), np.full(6, 1.1))
    ms = lmm.marginals(fx)
    assert [d.mu for d in ms] == [30.0] * 3 + [10.0] * 3 and approx([d.sigma for d in ms], np.full(6, math.sqrt(1.1)))
    assert len(lmm.rand(rng, fx)) == 2 * 3
    pfx = lmm.posterior(fx, y_train)
    p1, p2 = o.gp_posterior(g1, x_train, 0.1, y_1[tr]), o.gp_posterior(g2, x_train, 0.1, y_2[tr])
    post_fx = pfx(O(x_test, 2), 0.1)
    assert approx(lmm.logpdf(post_fx, y_test), o.finite_logpdf(p1, x_test, 0.1, y_1[te]) + o.finite_logpdf(p2, x_test, 0.1, y_2[te]))
    Mr = np.concatenate([o.gp_mean(p1, x_test), o.gp_mean(p2, x_test)])
    Vr = np.concatenate([o.gp_var(p1, x_test), o.gp_var(p2, x_test)]) + 0.1
    assert approx(lmm.mean(post_fx), Mr) and approx(lmm.var(post_fx), Vr)
    ms = lmm.marginals(post_fx)
    assert approx([d.mu for d in ms], Mr) and approx([d.sigma for d in ms], np.sqrt(Vr))
    assert len(lmm.rand(rng, post_fx)) == 2 * 2
    check_sampling_consistency(lmm, rng, f, O(x_train, 2))
    assert isinstance(lmm.logpdf_and_gradient(fx, y_train, with_grad_y=True), tuple)
    assert isinstance(lmm.logpdf_and_gradient(post_fx, y_test, with_grad_y=True), tuple)
    check_public_interface(lmm, rng, fx, secondary=False)
    check_public_interface(lmm, rng, post_fx, secondary=False)
    A = rng.standard_normal((6, 6))
    check_public_interface(lmm, rng, f(O(x_train, 2), A.T @ A + np.eye(6)))  # test/independent_mogp.jl:72-74


def approx_equivalent_to_dense(lmm, rng, fx, ofs, xpts, Sigma_feat, idx_of):
    """test/test_utils.jl:50-60 `approx_equivalent(rng, fx, fx_naive)`, fx_naive = the dense GP with
    LinearMixingModelKernel(kernels, I) evaluated by the oracle (by outputs) and reordered by features."""
    m = len(ofs)
    n = m * len(xpts)
    Cd = o.dense_mogp_cov(ofs, np.eye(m), xpts)[np.ix_(idx_of, idx_of)] + Sigma_feat
    Md = o.dense_mogp_mean(ofs, np.eye(m), xpts)[idx_of]
    assert len(fx) == n
    assert approx(lmm.mean(fx), Md, atol=1e-14)
    assert approx(lmm.var(fx), np.diag(Cd))
    assert approx(lmm.cov(fx), Cd)
    y = lmm.rand(rng, fx)
    L = np.linalg.cholesky(Cd)
    z = np.linalg.solve(L, y - Md)
    assert approx(lmm.logpdf(fx, y), -0.5 * (n * o.LOG2PI + 2 * np.sum(np.log(np.diag(L))) + z @ z))
    return y, Cd, Md


def test_independent_mogp_by_features(lmm):
    """test/independent_mogp.jl:80-145 "MOInputIsotopicByFeatures"."""
    F, O = lmm.MOInputIsotopicByFeatures, lmm.MOInputIsotopicByOutputs
    xpts = np.linspace(0.0, 2.0, 3)
    x = F(xpts, 2)
    # indices for reordering: the specific case where we know the answer
    v_by_output, v_by_features = np.array([1, 1, 1, 2, 2, 2]), np.array([1, 2, 1, 2, 1, 2])
    i_of = lmm.indices_which_reorder_outputs_to_features(x)
    i_fo = lmm.indices_which_reorder_features_to_outputs(x)
    assert np.array_equal(v_by_output[i_of], v_by_features) and np.array_equal(v_by_features[i_fo], v_by_output)
    assert np.array_equal(v_by_output[i_of][i_fo], v_by_output) and np.array_equal(v_by_features[i_fo][i_of], v_by_features)
    rng = np.random.default_rng(123456)
    ofs = [o.GP(o.Kernel(o.SE)), o.GP(o.Kernel(o.SE, 0.5))]
    f = lmm.IndependentMOGP([to_gp(lmm, g) for g in ofs])
    n = len(x)
    A = rng.standard_normal((n, n))
    for Sy in (0.1, np.ones(n) + rng.uniform(0, 1, n), A.T @ A + np.eye(n)):
        Sfeat = Sy * np.eye(n) if np.isscalar(Sy) else (np.diag(Sy) if Sy.ndim == 1 else Sy)
        fx = f(x, Sy)
        y, Cd, Md = approx_equivalent_to_dense(lmm, rng, fx, ofs, xpts, Sfeat, i_of)
        # posterior(fx, y)(x, Σy) against the dense conditional
        pf = lmm.posterior(fx, y)(x, Sy)
        Kxx = Cd - Sfeat
        Lc = np.linalg.cholesky(Cd)
        W = np.linalg.solve(Lc, Kxx)
        Mp = Md + Kxx @ np.linalg.solve(Cd, y - Md)
        Cp = Kxx - W.T @ W + Sfeat
        assert approx(lmm.mean(pf), Mp) and approx(lmm.var(pf), np.diag(Cp)) and approx(lmm.cov(pf), Cp)
        yp = lmm.rand(rng, pf)
        Lp = np.linalg.cholesky(Cp)
        zp = np.linalg.solve(Lp, yp - Mp)
        assert approx(lmm.logpdf(pf, yp), -0.5 * (n * o.LOG2PI + 2 * np.sum(np.log(np.diag(Lp))) + zp @ zp))
        check_public_interface(lmm, rng, fx)
    # mix of by-features and by-outputs: process covariances (test_internal_abstractgps_interface + cov checks)
    xq = np.linspace(0.0, 3.0, 4)
    xo = O(xq, 2)
    Kd = o.dense_mogp_cov(ofs, np.eye(2), xpts, xq)  # by outputs x by outputs
    assert approx(lmm.cov(f, x, xo), Kd[i_of, :]) and approx(lmm.cov(f, xo, x), Kd.T[:, i_of])
    assert approx(lmm.cov(f, x, xo), lmm.cov(f, xo, x).T)
    assert approx(np.diag(lmm.cov(f, x)), lmm.var(f(x, 0.0)))
    assert approx(lmm.cov(f, x), lmm.cov(f, x, x))
