"""Pins the CPU oracle with the reference's own equivalence identities and known answers
(SURVEY.md §4 / §8c): test/oilmm.jl, test/ilmm.jl, test/independent_mogp.jl, test/orthogonal_matrix.jl.
The reference holds no golden logpdf/mean/var numbers, so these identities are the pin."""
import math

import numpy as np
import pytest

from oracle import lmm_oracle as o


def toy_data(seed=4161999):
    """Shapes of generate_toy_data (test/test_utils.jl:1-35): 5 points on [0,10], 3/2 split, p=3."""
    rng = np.random.default_rng(seed)
    x = np.linspace(0.0, 10.0, 5)
    K = o.kernelmatrix(o.Kernel(o.SE), x) + 1e-6 * np.eye(5)
    ys = np.linalg.cholesky(K) @ rng.standard_normal((5, 3))
    idx = rng.permutation(5)
    tr, te = idx[:3], idx[3:]
    return x[tr], x[te], ys[tr].T.reshape(-1), ys[te].T.reshape(-1)


SHAPES = [
    (3, [o.SE, o.MATERN32, o.MATERN32]),  # test/oilmm.jl:44-49
    (2, [o.SE, o.MATERN32]),  # :51-56
    (1, [o.SE]),  # :58-63
]


@pytest.mark.parametrize("m,kinds", SHAPES)
def test_oilmm_equals_ilmm_and_dense(m, kinds):
    """test/oilmm.jl:10-27 (OILMM == ILMM(collect(H))) and test/ilmm.jl:10-27 (== dense MOGP)."""
    xtr, xte, ytr, yte = toy_data()
    U, S = o.orthogonal_from_seed(3, m, seed=7)
    fs = [o.GP(o.Kernel(k)) for k in kinds]
    model = o.OILMMModel(fs, U, S)
    H = model.H
    s2 = 0.1
    lo = o.oilmm_logpdf(model, xtr, s2, ytr)
    li = o.ilmm_logpdf(fs, H, xtr, s2, ytr)
    ld = o.dense_mogp_logpdf(fs, H, xtr, s2, ytr)
    assert lo == pytest.approx(ld, rel=1e-11)  # the regulariser restores the p-m orthogonal directions
    assert lo == pytest.approx(li, rel=1e-7)  # 1e-9 jitter in src/ilmm.jl:63
    # prior marginals
    Mo, Vo = o.oilmm_mean_and_var(model, xtr, s2)
    Mi, Vi = o.ilmm_mean_and_var(fs, H, xtr, s2)
    np.testing.assert_allclose(Mo, Mi, atol=1e-14)
    np.testing.assert_allclose(Vo, Vi, rtol=1e-12)
    # posterior marginals at test points
    post_o = o.oilmm_posterior(model, xtr, s2, ytr)
    post_i = o.ilmm_posterior(fs, H, xtr, s2, ytr)
    Mo, Vo = o.oilmm_mean_and_var(post_o, xte, s2)
    Mi, Vi = o.ilmm_mean_and_var(post_i, H, xte, s2)
    np.testing.assert_allclose(Mo, Mi, rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(Vo, Vi, rtol=1e-6)
    Md, Vd = o.dense_mogp_posterior_mean_and_var(fs, H, xtr, s2, ytr, xte, s2)
    np.testing.assert_allclose(Mo, Md, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(Vo, Vd, rtol=1e-9)


def test_oilmm_dense_identity_mid_size():
    """OILMM == dense pN x pN MVN when H is square (p == m); p > m differs only by
    the projection, so use the ILMM identity there."""
    rng = np.random.default_rng(0)
    N, p = 40, 4
    x = np.sort(rng.uniform(0, 4, N))
    U, S = o.orthogonal_from_seed(p, p, seed=3)
    fs = [o.GP(o.Kernel(k, 1.0, s)) for k, s in zip([o.SE, o.MATERN52, o.MATERN32, o.SE], [0.7, 1.3, 1.0, 1.9])]
    model = o.OILMMModel(fs, U, S)
    y = rng.standard_normal(N * p)
    assert o.oilmm_logpdf(model, x, 0.1, y) == pytest.approx(o.dense_mogp_logpdf(fs, model.H, x, 0.1, y), rel=1e-12)
    # p > m
    U, S = o.orthogonal_from_seed(5, 3, seed=3)
    model = o.OILMMModel(fs[:3], U, S)
    y = rng.standard_normal(N * 5)
    assert o.oilmm_logpdf(model, x, 0.1, y) == pytest.approx(o.ilmm_logpdf(fs[:3], model.H, x, 0.1, y), rel=1e-8)
    assert o.oilmm_logpdf(model, x, 0.1, y) == pytest.approx(o.dense_mogp_logpdf(fs[:3], model.H, x, 0.1, y), rel=1e-11)


def test_independent_mogp_equals_sum_of_gps():
    """test/independent_mogp.jl:33-60 with const means 30 and 10."""
    rng = np.random.default_rng(1)
    x = np.sort(rng.uniform(0, 3, 7))
    fs = [o.GP(o.Kernel(o.MATERN32), 30.0), o.GP(o.Kernel(o.SE), 10.0)]
    y = np.concatenate([30 + rng.standard_normal(7), 10 + rng.standard_normal(7)])
    s2 = 0.1
    total = o.imogp_logpdf(fs, x, s2, y)
    parts = o.gp_logpdf(fs[0], x, s2, y[:7]) + o.gp_logpdf(fs[1], x, s2, y[7:])
    assert total == pytest.approx(parts, rel=1e-15)
    # == dense LinearMixingModelKernel with H = I (test/independent_mogp.jl:112-128), scaled kernel
    fs2 = [o.GP(o.Kernel(o.SE, 0.5)), o.GP(o.Kernel(o.MATERN32))]
    yy = rng.standard_normal(14)
    assert o.imogp_logpdf(fs2, x, s2, yy) == pytest.approx(o.dense_mogp_logpdf(fs2, np.eye(2), x, s2, yy), rel=1e-12)
    M, V = o.imogp_mean_and_var(fs, x, s2)
    np.testing.assert_allclose(M, np.r_[np.full(7, 30.0), np.full(7, 10.0)])
    np.testing.assert_allclose(V, 1.0 + s2)
    # posterior of each output = single-GP posterior
    posts = o.imogp_posterior(fs, x, s2, y)
    single = o.gp_posterior(fs[1], x, s2, y[7:])
    np.testing.assert_array_equal(posts[1].alpha, single.alpha)


def test_by_features_permutation_known_answers():
    """test/independent_mogp.jl:86-98: [1,1,1,2,2,2] <-> [1,2,1,2,1,2]."""
    by_out = np.array([1, 1, 1, 2, 2, 2])
    by_feat = np.array([1, 2, 1, 2, 1, 2])
    N, p = 3, 2
    np.testing.assert_array_equal(by_out[o.indices_outputs_to_features(N, p)], by_feat)
    np.testing.assert_array_equal(by_feat[o.indices_features_to_outputs(N, p)], by_out)
    # 1-based known index vectors from Julia: vec(reshape(1:6,3,2)') = [1,4,2,5,3,6]
    np.testing.assert_array_equal(o.indices_outputs_to_features(3, 2) + 1, [1, 4, 2, 5, 3, 6])
    np.testing.assert_array_equal(o.indices_features_to_outputs(3, 2) + 1, [1, 3, 5, 2, 4, 6])
    rng = np.random.default_rng(2)
    v = rng.standard_normal(12)
    np.testing.assert_array_equal(v[o.indices_outputs_to_features(4, 3)][o.indices_features_to_outputs(4, 3)], v)


def test_reshape_y_and_unpack_known_answers():
    """test/ilmm.jl:55-68."""
    y = np.arange(16.0)
    assert o.reshape_y(y, 8).shape == (2, 8)
    assert o.reshape_y(y, 2).shape == (8, 2)
    assert o.reshape_y(y, 8)[1, 0] == 8.0  # Y[j,i] = y[(j-1)N+i]
    fs = [o.GP()]
    with pytest.raises(RuntimeError, match="out dim of x != out dim of f."):
        o.oilmm_logpdf(o.OILMMModel(fs, np.ones((3, 1)) / math.sqrt(3), np.ones(1)), np.arange(2.0), 0.1, np.zeros(4))


def test_orthogonal_validation():
    """test/orthogonal_matrix.jl:1-15."""
    rng = np.random.default_rng(3)
    with pytest.raises(ValueError, match="not an orthogonal matrix"):
        o.validate_orthogonal(rng.uniform(size=(3, 2)))
    U, S = o.orthogonal_from_seed(3, 2)
    o.validate_orthogonal(U)
    m = o.OILMMModel([], U, S)
    np.testing.assert_allclose(m.H, U @ np.diag(np.sqrt(S)))


def test_distance_formulations_spread():
    """gemm-trick vs direct distances agree far below 1e-9 when well conditioned."""
    rng = np.random.default_rng(5)
    N = 400
    x = np.sort(rng.uniform(0, N / 100, N))
    for kind in (o.SE, o.MATERN52):
        f = o.GP(o.Kernel(kind, 1.0, 1.3))
        y = rng.standard_normal(N)
        a = o.gp_logpdf(f, x, 0.05, y, form="gemm")
        b = o.gp_logpdf(f, x, 0.05, y, form="direct")
        assert a == pytest.approx(b, rel=1e-11)


def test_sampling_consistency():
    """test/test_utils.jl:41-48: sample at σ²=1e-6, condition, posterior mean ≈ y (rtol 1e-2)."""
    rng = np.random.default_rng(6)
    x = np.array([0.0, 2.5, 7.5])
    U, S = o.orthogonal_from_seed(3, 2, seed=9)
    model = o.OILMMModel([o.GP(o.Kernel(o.SE)), o.GP(o.Kernel(o.MATERN32))], U, S)
    # sample from the *projected* model so that y lies in span(H) up to noise
    y = o.oilmm_rand(model, x, 1e-6, rng.standard_normal(6), rng.standard_normal(9))
    post = o.oilmm_posterior(model, x, 1e-6, y)
    M, V = o.oilmm_mean_and_var(post, x, 1e-6)
    np.testing.assert_allclose(M, y, rtol=1e-2, atol=1e-2)
    assert np.all(V < 1e-3)


def test_multidim_inputs_colvecs():
    """x as ColVecs / RowVecs (test/ilmm.jl:65): D-dimensional inputs go through the same path."""
    rng = np.random.default_rng(8)
    X = rng.uniform(0, 2, size=(30, 3))
    f = o.GP(o.Kernel(o.MATERN52, 0.8, 0.9))
    y = rng.standard_normal(30)
    a = o.gp_logpdf(f, X, 0.1, y, form="gemm")
    b = o.gp_logpdf(f, X, 0.1, y, form="direct")
    assert a == pytest.approx(b, rel=1e-12)


def test_oracle_gradient_matches_finite_differences():
    """The analytic logpdf gradient (used to check the CUDA rrule) against central differences."""
    rng = np.random.default_rng(9)
    N, p, m = 25, 4, 3
    x = np.sort(rng.uniform(0, 4, N))
    U, S = o.orthogonal_from_seed(p, m, seed=5)
    fs = [o.GP(o.Kernel(o.SE, 0.9, 1.2), 0.3), o.GP(o.Kernel(o.MATERN32, 1.3, 0.8), -0.2), o.GP(o.Kernel(o.MATERN52, 0.7, 1.5), 0.1)]
    y = rng.standard_normal(p * N)
    lp, g = o.oilmm_logpdf_grad(o.OILMMModel(fs, U, S), x, 0.2, y)
    assert lp == pytest.approx(o.oilmm_logpdf(o.OILMMModel(fs, U, S), x, 0.2, y), rel=1e-12)

    def f_at(i, field, h):
        k = fs[i].kernel
        kw = dict(kind=k.kind, variance=k.variance, inv_lengthscale=k.inv_lengthscale)
        mean = fs[i].mean_const
        if field == "mean_const":
            mean += h
        else:
            kw[field] += h
        fs2 = list(fs)
        fs2[i] = o.GP(o.Kernel(**kw), mean)
        return o.oilmm_logpdf(o.OILMMModel(fs2, U, S), x, 0.2, y, form="direct")

    h = 1e-6
    for i in range(m):
        for field in ("variance", "inv_lengthscale", "mean_const"):
            fd = (f_at(i, field, h) - f_at(i, field, -h)) / (2 * h)
            assert g[field][i] == pytest.approx(fd, rel=2e-6, abs=1e-7)
    model = o.OILMMModel(fs, U, S)
    fd = (o.oilmm_logpdf(model, x, 0.2 + h, y) - o.oilmm_logpdf(model, x, 0.2 - h, y)) / (2 * h)
    assert g["sigma2"] == pytest.approx(fd, rel=2e-6)
    for j in (0, 17, p * N - 1):
        e = np.zeros(p * N)
        e[j] = h
        fd = (o.oilmm_logpdf(model, x, 0.2, y + e) - o.oilmm_logpdf(model, x, 0.2, y - e)) / (2 * h)
        assert g["y"][j] == pytest.approx(fd, rel=2e-6, abs=1e-8)
    # mixing matrix fields (unconstrained Euclidean derivatives of the reference's expressions)
    for i in range(m):
        Sp, Sm = S.copy(), S.copy()
        Sp[i] += h
        Sm[i] -= h
        fd = (o.oilmm_logpdf(o.OILMMModel(fs, U, Sp), x, 0.2, y) - o.oilmm_logpdf(o.OILMMModel(fs, U, Sm), x, 0.2, y)) / (2 * h)
        assert g["S"][i] == pytest.approx(fd, rel=2e-6)
    for (j, i) in ((0, 0), (2, 1), (3, 2)):
        Up, Um = U.copy(), U.copy()
        Up[j, i] += h
        Um[j, i] -= h
        fd = (o.oilmm_logpdf(o.OILMMModel(fs, Up, S), x, 0.2, y) - o.oilmm_logpdf(o.OILMMModel(fs, Um, S), x, 0.2, y)) / (2 * h)
        assert g["U"][j, i] == pytest.approx(fd, rel=5e-6, abs=1e-6)


def test_oracle_ilmm_gradient_matches_finite_differences():
    """Analytic gradient of the general-ILMM logpdf (src/ilmm.jl:150-163 through `project` and
    `regulariser`; the oracle for lmm_ilmm_logpdf_grad) against central differences."""
    rng = np.random.default_rng(3)
    N, p, m = 17, 4, 3
    x = np.sort(rng.uniform(0, 4, N))
    H = rng.uniform(0, 1, (p, m))
    y = rng.standard_normal(p * N)
    s2 = 0.3
    par = np.array([[1.3, 0.7, 0.2], [0.8, 1.4, -0.5], [1.1, 0.9, 1.0]])
    kinds = [o.SE, o.MATERN32, o.MATERN52]

    def mk(q):
        return [o.GP(o.Kernel(kinds[i], q[i, 0], q[i, 1]), q[i, 2]) for i in range(m)]

    def f(q=par, Hm=H, s=s2, yy=y):
        return o.ilmm_logpdf(mk(q), Hm, x, s, yy, form="direct")

    lp, g = o.ilmm_logpdf_grad(mk(par), H, x, s2, y)
    assert lp == pytest.approx(f(), rel=1e-12)
    h = 1e-6
    for i in range(m):
        for c, field in enumerate(("variance", "inv_lengthscale", "mean_const")):
            qp, qm = par.copy(), par.copy()
            qp[i, c] += h
            qm[i, c] -= h
            assert g[field][i] == pytest.approx((f(q=qp) - f(q=qm)) / (2 * h), rel=5e-6, abs=1e-6)
    assert g["sigma2"] == pytest.approx((f(s=s2 + h) - f(s=s2 - h)) / (2 * h), rel=5e-6)
    for j in (0, 23, p * N - 1):
        e = np.zeros(p * N)
        e[j] = h
        assert g["y"][j] == pytest.approx((f(yy=y + e) - f(yy=y - e)) / (2 * h), rel=5e-6, abs=1e-7)
    for j in range(p):
        for i in range(m):
            Hp, Hm_ = H.copy(), H.copy()
            Hp[j, i] += h
            Hm_[j, i] -= h
            assert g["H"][j, i] == pytest.approx((f(Hm=Hp) - f(Hm=Hm_)) / (2 * h), rel=5e-6, abs=1e-6)


def test_oracle_posterior_logpdf_gradient_matches_finite_differences():
    """d/dσ² and d/dy* of logpdf(post(x*, σ²), y*) for OILMM, IndependentMOGP and general-ILMM posteriors."""
    rng = np.random.default_rng(5)
    N, p, m, Ns = 20, 4, 3, 6
    x = np.sort(rng.uniform(0, 4, N))
    xs = rng.uniform(0, 4, Ns)
    U, S = o.orthogonal_from_seed(p, m, seed=2)
    fs = [o.GP(o.Kernel(o.SE, 0.9, 1.2), 0.3), o.GP(o.Kernel(o.MATERN32, 1.3, 0.8), -0.2), o.GP(o.Kernel(o.MATERN52, 0.7, 1.5), 0.1)]
    y, ys = rng.standard_normal(p * N), rng.standard_normal(p * Ns)
    h = 1e-6
    e = np.zeros(p * Ns)
    e[7] = h
    post = o.oilmm_posterior(o.OILMMModel(fs, U, S), x, 0.1, y)
    lp, g = o.oilmm_post_logpdf_grad(post, xs, 0.2, ys)
    assert lp == pytest.approx(o.oilmm_logpdf(post, xs, 0.2, ys), rel=1e-12)
    assert g["sigma2"] == pytest.approx((o.oilmm_logpdf(post, xs, 0.2 + h, ys) - o.oilmm_logpdf(post, xs, 0.2 - h, ys)) / (2 * h), rel=5e-6)
    assert g["y"][7] == pytest.approx((o.oilmm_logpdf(post, xs, 0.2, ys + e) - o.oilmm_logpdf(post, xs, 0.2, ys - e)) / (2 * h), rel=5e-6)
    posts = o.imogp_posterior(fs, x, 0.1, y[: m * N])
    yi = ys[: m * Ns]
    ref = lambda s2: sum(o.finite_logpdf(posts[i], xs, s2, yi.reshape(m, Ns)[i]) for i in range(m))
    lp, g = o.oilmm_post_logpdf_grad(o.OILMMModel(posts, np.eye(m), np.ones(m)), xs, 0.2, yi)
    assert lp == pytest.approx(ref(0.2), rel=1e-12)
    assert g["sigma2"] == pytest.approx((ref(0.2 + h) - ref(0.2 - h)) / (2 * h), rel=5e-6)
    H = rng.uniform(0, 1, (p, m))
    ip = o.ilmm_posterior(fs, H, x, 0.1, y)
    lp, g = o.ilmm_post_logpdf_grad(ip, xs, 0.2, ys)
    assert lp == pytest.approx(o.ilmm_post_logpdf(ip, xs, 0.2, ys), rel=1e-12)
    assert g["sigma2"] == pytest.approx((o.ilmm_post_logpdf(ip, xs, 0.2 + h, ys) - o.ilmm_post_logpdf(ip, xs, 0.2 - h, ys)) / (2 * h), rel=5e-6)
    assert g["y"][7] == pytest.approx((o.ilmm_post_logpdf(ip, xs, 0.2, ys + e) - o.ilmm_post_logpdf(ip, xs, 0.2, ys - e)) / (2 * h), rel=5e-6)


def test_oracle_exponential_and_rational_quadratic_kernels():
    """ExponentialKernel (= Matern12Kernel, exp(-d)) and RationalQuadraticKernel(α) ((1 + d²/2α)^(-α)), with an ARDTransform:
    known values, and the lengthscale derivative against central differences."""
    x = np.array([[0.0, 0.0], [3.0, 4.0]])
    assert o.kernelmatrix(o.Kernel(o.EXPONENTIAL), x)[0, 1] == pytest.approx(math.exp(-5.0), rel=1e-15)
    assert o.kernelmatrix(o.Kernel(o.RATQUAD, param=2.0), x)[0, 1] == pytest.approx((1 + 25.0 / 4.0) ** -2.0, rel=1e-15)
    assert o.kernelmatrix(o.Kernel(o.EXPONENTIAL, 2.0, 0.5, (2.0, 1.0)), x)[0, 1] == pytest.approx(2.0 * math.exp(-math.sqrt(9.0 + 4.0)), rel=1e-15)
    rng = np.random.default_rng(4)
    X = rng.uniform(0, 2, (12, 2))
    y = rng.standard_normal(12)
    h = 1e-6
    for k in (o.Kernel(o.EXPONENTIAL, 0.8, 1.3, (0.7, 1.4)), o.Kernel(o.RATQUAD, 1.2, 0.9, None, 1.7)):
        _, _, gs, _, _, _ = o.gp_logpdf_grad(o.GP(k, 0.1), X, 0.2, y)
        kp = o.Kernel(k.kind, k.variance, k.inv_lengthscale + h, k.ard, k.param)
        km = o.Kernel(k.kind, k.variance, k.inv_lengthscale - h, k.ard, k.param)
        fd = (o.gp_logpdf(o.GP(kp, 0.1), X, 0.2, y, form="direct") - o.gp_logpdf(o.GP(km, 0.1), X, 0.2, y, form="direct")) / (2 * h)
        assert gs == pytest.approx(fd, rel=2e-6, abs=1e-8)


def test_oracle_ard_gradient_matches_finite_differences():
    rng = np.random.default_rng(11)
    X = rng.uniform(0, 2, (14, 3))
    y = rng.standard_normal(14)
    h = 1e-6
    for kind, param in ((o.SE, 1.0), (o.MATERN32, 1.0), (o.MATERN52, 1.0), (o.EXPONENTIAL, 1.0), (o.RATQUAD, 1.4)):
        ard = (0.7, 1.3, 2.1)
        g = o.gp_logpdf_grad_ard(o.GP(o.Kernel(kind, 0.9, 1.1, ard, param), 0.2), X, 0.15, y)
        for j in range(3):
            ap, am = list(ard), list(ard)
            ap[j] += h
            am[j] -= h
            fd = (o.gp_logpdf(o.GP(o.Kernel(kind, 0.9, 1.1, tuple(ap), param), 0.2), X, 0.15, y, form="direct")
                  - o.gp_logpdf(o.GP(o.Kernel(kind, 0.9, 1.1, tuple(am), param), 0.2), X, 0.15, y, form="direct")) / (2 * h)
            assert g[j] == pytest.approx(fd, rel=5e-6, abs=1e-8)
