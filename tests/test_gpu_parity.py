"""GPU parity tests proper: the CUDA path (through the C ABI / host mirror) against the CPU
oracle on identical seeded inputs.  Tolerance: 1e-9 relative on logpdf, posterior means and
variances (BASELINE.json north_star), written next to each assertion."""
import numpy as np
import pytest
import scipy.linalg as sla

from oracle import lmm_oracle as o
from _tol import assert_isapprox, relnorm

pytestmark = pytest.mark.gpu

RTOL = 1e-9


@pytest.fixture(scope="module")
def lmm():
    import lmm_b200

    lmm_b200.default_context()  # fails loudly without a GPU / without liblmm.so
    return lmm_b200


KMAP = {o.SE: "SEKernel", o.MATERN32: "Matern32Kernel", o.MATERN52: "Matern52Kernel", o.EXPONENTIAL: "ExponentialKernel"}


def to_lmm_gp(lmm, g: o.GP):
    k = lmm.RationalQuadraticKernel(g.kernel.param) if g.kernel.kind == o.RATQUAD else getattr(lmm, KMAP[g.kernel.kind])()
    k = g.kernel.variance * k if g.kernel.variance != 1.0 else k
    if g.kernel.inv_lengthscale != 1.0:
        k = k.compose(lmm.ScaleTransform(g.kernel.inv_lengthscale))
    if g.kernel.ard is not None:
        k = k.compose(lmm.ARDTransform(g.kernel.ard))
    return lmm.GP(g.mean_const, k)


def make_problem(N, p, m, Ns, seed=0, D=1, means=False, kinds=None):
    rng = np.random.default_rng(seed)
    if D == 1:
        x = np.sort(rng.uniform(0, max(N / 100.0, 4.0), N))
        xs = rng.uniform(0, max(N / 100.0, 4.0), Ns)
    else:
        x = rng.uniform(0, 3.0, (N, D))
        xs = rng.uniform(0, 3.0, (Ns, D))
    U, S = o.orthogonal_from_seed(p, m, seed=seed + 1)
    kinds = kinds or [o.SE, o.MATERN32, o.MATERN52]
    fs = [
        o.GP(o.Kernel(kinds[i % len(kinds)], float(rng.uniform(0.5, 1.5)), float(rng.uniform(0.5, 2.0))),
             float(rng.normal()) if means else 0.0)
        for i in range(m)
    ]
    y = rng.standard_normal(p * N)
    return x, xs, U, S, fs, y


def rel(a, b):
    return abs(a - b) / max(abs(b), 1e-300)


@pytest.mark.parametrize("N,batch", [(50, 2), (128, 1), (300, 3), (1000, 2), (2000, 1), (1700, 2)])
def test_potrf_batched_matches_lapack(lmm, N, batch):
    rng = np.random.default_rng(N)
    A = rng.standard_normal((batch, N, N))
    A = A @ np.transpose(A, (0, 2, 1)) / N + np.eye(N)[None] * 0.5
    L, logdet, info = lmm.potrf_batched(A)
    assert np.all(info == 0)
    for b in range(batch):
        Lr = sla.cholesky(A[b], lower=True)
        np.testing.assert_allclose(L[b], Lr, rtol=1e-10, atol=1e-12)
        assert rel(logdet[b], 2 * np.sum(np.log(np.diag(Lr)))) < 1e-11 or abs(logdet[b]) < 1e-9
        assert np.all(np.triu(L[b], 1) == 0.0)


def test_potrf_small_batch_schedules(lmm):
    """Batch-1 / batch-2 schedules -- plain, right-looking look-ahead on two streams, each with the panel chain as one fused
    launch per tile column (chain_column_kernel: TRSM + next-column update + next diagonal tile handed over through ready
    counters) or as three PDL-chained launches -- at several block widths all reproduce LAPACK's factor.  N = 2300 (18 tile
    columns: block boundaries, a ragged last tile) and N = 4500 (36 columns: automatic block width 2)."""
    ctx = lmm.default_context()
    try:
        for N, batch in ((2300, 1), (2300, 2), (4500, 1)):
            rng = np.random.default_rng(N)
            A = rng.standard_normal((batch, N, N))
            A = A @ np.transpose(A, (0, 2, 1)) / N + np.eye(N)[None]
            Lr = [sla.cholesky(A[b], lower=True) for b in range(batch)]
            configs = ((0, 0, 0, 0), (0, 0, 1, 1), (1, 0, 0, 0), (1, 1, 0, 1), (1, 3, 0, 1), (1, 0, 1, 0), (1, 0, 1, 1), (1, 1, 1, 1), (1, 2, 1, 1),
                       (1, 3, 1, 1), (1, 5, 1, 0), (1, 7, 1, 1), (1, 18, 1, 1))
            for la, ob, fused, pdl in (configs if N == 2300 else ((1, 0, 2, 1), (1, 4, 2, 1), (1, 0, 1, 1), (1, 0, 0, 1))):
                ctx.set_option("lookahead", la)
                ctx.set_option("outer_block", ob)
                ctx.set_option("chain_fused", fused)
                ctx.set_option("pdl", pdl)
                L, logdet, info = lmm.potrf_batched(A)
                for b in range(batch):
                    assert info[b] == 0
                    np.testing.assert_allclose(L[b], Lr[b], rtol=1e-10, atol=1e-12, err_msg=f"N={N} batch={batch} lookahead={la} outer_block={ob} chain_fused={fused} pdl={pdl}")
                    assert rel(logdet[b], 2 * np.sum(np.log(np.diag(Lr[b])))) < 1e-11
    finally:
        ctx.set_option("lookahead", 1)
        ctx.set_option("outer_block", 0)
        ctx.set_option("chain_fused", 1)
        ctx.set_option("pdl", 1)


@pytest.mark.parametrize("fused", [0, 1])
def test_potrf_reports_non_pd(lmm, fused):
    """LAPACK `info` semantics (1-based index of the first non-positive pivot), through the standalone diagonal-tile kernel and
    through the fused chain kernel's factorisation role; the pivot sits in the second tile and in the middle of an 8-column
    panel step."""
    A = np.eye(200)
    A[150, 150] = -1.0
    B = np.eye(200)
    B[5, 5] = 0.0
    ctx = lmm.default_context()
    ctx.set_option("chain_fused", fused)
    try:
        L, logdet, info = lmm.potrf_batched(np.stack([np.eye(200), A, B]))
    finally:
        ctx.set_option("chain_fused", 1)
    assert info[0] == 0 and info[1] == 151 and info[2] == 6


@pytest.mark.parametrize(
    "N,p,m,Ns,D,means",
    [(50, 3, 2, 7, 1, False), (3, 3, 3, 2, 1, False), (700, 8, 4, 33, 1, True), (260, 5, 5, 130, 2, True), (1, 2, 1, 1, 1, False)],
)
def test_oilmm_logpdf_posterior_marginals(lmm, N, p, m, Ns, D, means):
    """src/oilmm.jl:79-93, 116-134, 57-76 vs oracle (BASELINE config 1 is the first case)."""
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=N, D=D, means=means)
    s2 = 0.1
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    assert isinstance(f, lmm.OILMM)
    xin = lmm.MOInputIsotopicByOutputs(x if D == 1 else lmm.RowVecs(x), p)
    xsin = lmm.MOInputIsotopicByOutputs(xs if D == 1 else lmm.RowVecs(xs), p)
    fx = f(xin, s2)
    ref_terms, ref_reg = o.oilmm_logpdf_terms(om, x, s2, y)
    lp = lmm.logpdf(fx, y)
    assert rel(lp, float(np.sum(ref_terms) + ref_reg)) < RTOL
    terms = lmm.logpdf_terms(fx, y)
    np.testing.assert_allclose(terms[:m], ref_terms, rtol=RTOL)
    assert rel(terms[m], ref_reg) < RTOL or abs(ref_reg) < 1e-9
    post, lp2 = lmm.posterior(fx, y, with_logpdf=True)
    assert lp2 == lp
    M, V = lmm.mean_and_var(post(xsin, s2))
    opost = o.oilmm_posterior(om, x, s2, y)
    Mr, Vr = o.oilmm_mean_and_var(opost, xs, s2)
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(V, Vr, rtol=RTOL)
    # PosteriorGP fields (α, C, δ)
    g0 = post.f.fs[m - 1]
    assert_isapprox(g0.alpha, opost.fs[m - 1].alpha, RTOL)
    np.testing.assert_allclose(g0.delta, opost.fs[m - 1].delta, rtol=1e-12, atol=1e-13)
    assert_isapprox(g0.C, opost.fs[m - 1].L, RTOL)
    # prior marginals
    Mp, Vp = lmm.mean_and_var(f(xsin, s2))
    Mpr, Vpr = o.oilmm_mean_and_var(om, xs, s2)
    assert_isapprox(Mp, Mpr, RTOL)
    np.testing.assert_allclose(Vp, Vpr, rtol=RTOL)
    marg = lmm.marginals(post(xsin, s2))
    assert len(marg) == p * Ns and abs(marg[0].sigma ** 2 - V[0]) < 1e-12


def test_oilmm_mid_size_multi_tile(lmm):
    """Several tile columns and more than one outer block: N = 1300 (11 tiles)."""
    N, p, m, Ns = 1300, 6, 3, 200
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=5, means=True)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    assert rel(lmm.logpdf(fx, y), o.oilmm_logpdf(om, x, 0.1, y)) < RTOL
    post = lmm.posterior(fx, y)
    M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
    Mr, Vr = o.oilmm_mean_and_var(o.oilmm_posterior(om, x, 0.1, y), xs, 0.1)
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(V, Vr, rtol=RTOL)


@pytest.mark.parametrize("streams,outer,small,fused", [(1, 8, 74, 1), (1, 3, 74, 0), (4, 16, 0, 1), (8, 1, 0, 0), (2, 8, 0, 1), (1, 5, 74, 1),
                                                       (4, 8, 4096, 1), (4, 8, 4096, 0), (1, 2, 4096, 1)])
def test_streams_blocking_and_small_grid_kernels_agree(lmm, streams, outer, small, fused):
    """The TMA-pipelined GEMM, the latency-optimised direct kernel for small grids (off / default threshold / everywhere), the
    fused per-column chain kernel, any stream-group count and any outer block width give the same factor: batched OILMM
    (N = 1100: 9 tile columns, batch 3 -- small enough for the fused chain) and a single general-ILMM factor (batch 1,
    right-looking schedule)."""
    N, p, m, Ns = 1100, 5, 3, 70
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=77, means=True)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    ctx = lmm.default_context()
    ctx.set_option("streams", streams)
    ctx.set_option("outer_block", outer)
    ctx.set_option("gemm_small", small)
    ctx.set_option("chain_fused", fused)
    try:
        fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
        post, lp = lmm.posterior(fx, y, with_logpdf=True)
        assert rel(lp, o.oilmm_logpdf(om, x, 0.1, y)) < RTOL
        M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
        Mr, Vr = o.oilmm_mean_and_var(o.oilmm_posterior(om, x, 0.1, y), xs, 0.1)
        assert_isapprox(M, Mr, RTOL)
        np.testing.assert_allclose(V, Vr, rtol=RTOL)
        assert_isapprox(post.f.fs[m - 1].C, o.oilmm_posterior(om, x, 0.1, y).fs[m - 1].L, RTOL, "factor")
        # general ILMM: one (mN x mN) factor, batch 1
        rng = np.random.default_rng(5)
        N2, p2, m2 = 600, 4, 3
        x2 = np.sort(rng.uniform(0, 6, N2))
        H = rng.uniform(0, 1, (p2, m2))
        fs2 = [o.GP(o.Kernel(k, 1.0, s)) for k, s in zip([o.SE, o.MATERN32, o.MATERN52], [0.9, 1.2, 1.5])]
        y2 = rng.standard_normal(p2 * N2)
        f2 = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs2]), H)
        assert rel(lmm.logpdf(f2(lmm.MOInputIsotopicByOutputs(x2, p2), 0.1), y2), o.ilmm_logpdf(fs2, H, x2, 0.1, y2)) < RTOL
    finally:
        ctx.set_option("streams", 4)
        ctx.set_option("outer_block", 0)
        ctx.set_option("gemm_small", 74)
        ctx.set_option("chain_fused", 1)


def test_distance_form_option(lmm):
    x, xs, U, S, fs, y = make_problem(300, 4, 2, 5, seed=11)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    fx = f(lmm.MOInputIsotopicByOutputs(x, 4), 0.1)
    ctx = lmm.default_context()
    ctx.set_option("distance_form", 1)
    try:
        assert rel(lmm.logpdf(fx, y), o.oilmm_logpdf(om, x, 0.1, y, form="direct")) < RTOL
    finally:
        ctx.set_option("distance_form", 0)


def test_independent_mogp(lmm):
    """src/independent_mogp.jl:74-80,119-126,222-229 with const means (test/independent_mogp.jl:33-34)."""
    rng = np.random.default_rng(3)
    N, Ns = 150, 11
    x = np.sort(rng.uniform(0, 5, N))
    xs = rng.uniform(0, 5, Ns)
    fs = [o.GP(o.Kernel(o.MATERN32), 30.0), o.GP(o.Kernel(o.SE, 0.5), 10.0)]
    y = np.concatenate([30 + rng.standard_normal(N), 10 + rng.standard_normal(N)])
    f = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    fx = f(lmm.MOInputIsotopicByOutputs(x, 2), 0.1)
    assert rel(lmm.logpdf(fx, y), o.imogp_logpdf(fs, x, 0.1, y)) < RTOL
    # by features
    yf = y[o.indices_outputs_to_features(N, 2)]
    fxf = f(lmm.MOInputIsotopicByFeatures(x, 2), 0.1)
    assert rel(lmm.logpdf(fxf, yf), o.imogp_logpdf(fs, x, 0.1, y)) < RTOL
    post = lmm.posterior(fx, y)
    M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, 2), 0.1))
    posts = o.imogp_posterior(fs, x, 0.1, y)
    Mr, Vr = o.imogp_mean_and_var(posts, xs, 0.1)
    np.testing.assert_allclose(M, Mr, rtol=RTOL)
    np.testing.assert_allclose(V, Vr, rtol=RTOL)
    Mf, Vf = lmm.mean_and_var(post(lmm.MOInputIsotopicByFeatures(xs, 2), 0.1))
    np.testing.assert_allclose(Mf, Mr[o.indices_outputs_to_features(Ns, 2)], rtol=RTOL)
    # rand: deterministic given the normals
    z = np.random.default_rng(9).standard_normal(2 * N)
    s = lmm.rand(np.random.default_rng(9), fx)
    np.testing.assert_allclose(s, o.imogp_rand(fs, x, 0.1, z), rtol=1e-8, atol=1e-9)


@pytest.mark.parametrize("form", [0, 1])
def test_ilmm_logpdf_forms(lmm, form):
    """src/ilmm.jl:150-163 (projected form) and the dense pN form (test/ilmm.jl:5)."""
    rng = np.random.default_rng(4)
    N, p, m = 200, 4, 3
    x = np.sort(rng.uniform(0, 4, N))
    H = rng.uniform(0, 1, (p, m))
    fs = [o.GP(o.Kernel(o.SE, 1.0, 1.1), 0.3), o.GP(o.Kernel(o.MATERN32)), o.GP(o.Kernel(o.MATERN52, 0.7, 0.8), -0.2)]
    y = rng.standard_normal(p * N)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), H)
    assert not isinstance(f, lmm.OILMM)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    lmm.set_ilmm_form(form)
    try:
        got = lmm.logpdf(fx, y)
    finally:
        lmm.set_ilmm_form(0)
    ref = o.ilmm_logpdf(fs, H, x, 0.1, y) if form == 0 else o.dense_mogp_logpdf(fs, H, x, 0.1, y)
    assert rel(got, ref) < RTOL


def test_ilmm_posterior_marginals(lmm):
    """src/ilmm.jl:184-198 + 122-129 on the joint posterior."""
    rng = np.random.default_rng(6)
    N, Ns, p, m = 150, 9, 3, 2
    x = np.sort(rng.uniform(0, 4, N))
    xs = rng.uniform(0, 4, Ns)
    H = rng.uniform(0, 1, (p, m))
    fs = [o.GP(o.Kernel(o.SE), 0.5), o.GP(o.Kernel(o.MATERN32, 0.8, 1.2))]
    y = rng.standard_normal(p * N)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), H)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    post, lp = lmm.posterior(fx, y, with_logpdf=True)
    assert rel(lp, o.ilmm_logpdf(fs, H, x, 0.1, y)) < RTOL
    M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
    Mr, Vr = o.ilmm_mean_and_var(o.ilmm_posterior(fs, H, x, 0.1, y), H, xs, 0.1)
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(V, Vr, rtol=RTOL)
    Mp, Vp = lmm.mean_and_var(f(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
    Mpr, Vpr = o.ilmm_mean_and_var(fs, H, xs, 0.1)
    assert_isapprox(Mp, Mpr, RTOL)
    np.testing.assert_allclose(Vp, Vpr, rtol=RTOL)


def test_rand_prior_and_posterior(lmm):
    """src/oilmm.jl:40-54 and src/ilmm.jl:78-87: samples are a deterministic function of the normals."""
    rng = np.random.default_rng(8)
    N, p, m, Ns = 6, 3, 2, 4
    x = np.linspace(0, 10, N)
    xs = np.array([1.3, 4.1, 6.2, 9.4])
    U, S = o.orthogonal_from_seed(p, m, seed=2)
    fs = [o.GP(o.Kernel(o.MATERN32)), o.GP(o.Kernel(o.MATERN32, 1.0, 2.0), 0.4)]
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    g = np.random.default_rng(21)
    zl, zn = g.standard_normal(m * N), g.standard_normal(p * N)
    s = lmm.rand(np.random.default_rng(21), fx)
    assert s.shape == (p * N,)
    np.testing.assert_allclose(s, o.oilmm_rand(om, x, 0.1, zl, zn), rtol=1e-7, atol=1e-8)
    Ys = lmm.rand(np.random.default_rng(1), fx, 3)
    assert Ys.shape == (p * N, 3)
    lps = lmm.logpdf(fx, Ys)  # AbstractGPs logpdf(fx, Y::AbstractMatrix): one value per column
    assert lps.shape == (3,) and rel(lps[1], o.oilmm_logpdf(om, x, 0.1, Ys[:, 1])) < RTOL
    # general ILMM
    Hm = om.H
    fi = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), Hm)
    si = lmm.rand(np.random.default_rng(21), fi(lmm.MOInputIsotopicByOutputs(x, p), 0.1))
    np.testing.assert_allclose(si, o.ilmm_rand(fs, Hm, x, 0.1, zl, zn), rtol=1e-7, atol=1e-8)
    # posterior sample and posterior logpdf at test points (test/oilmm.jl:84-86)
    y = o.oilmm_rand(om, x, 0.1, zl, zn)
    post = lmm.posterior(fx, y)
    opost = o.oilmm_posterior(om, x, 0.1, y)
    g = np.random.default_rng(22)
    zl2, zn2 = g.standard_normal(m * Ns), g.standard_normal(p * Ns)
    pfx = post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1)
    sp = lmm.rand(np.random.default_rng(22), pfx)
    np.testing.assert_allclose(sp, o.oilmm_rand(opost, xs, 0.1, zl2, zn2), rtol=1e-6, atol=1e-7)
    ys = np.random.default_rng(23).standard_normal(p * Ns)
    assert rel(lmm.logpdf(pfx, ys), o.oilmm_logpdf(opost, xs, 0.1, ys)) < RTOL


def test_posterior_logpdf_larger(lmm):
    N, p, m, Ns = 400, 4, 3, 150
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=31, means=True)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    post = lmm.posterior(f(lmm.MOInputIsotopicByOutputs(x, p), 0.1), y)
    ys = np.random.default_rng(1).standard_normal(p * Ns)
    got = lmm.logpdf(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.2), ys)
    ref = o.oilmm_logpdf(o.oilmm_posterior(om, x, 0.1, y), xs, 0.2, ys)
    assert rel(got, ref) < RTOL


def test_logpdf_sweep(lmm):
    """BASELINE config 5 shape at test scale: one call, several lengthscale settings."""
    N, p, m = 300, 6, 3
    x, _, U, S, fs, y = make_problem(N, p, m, 1, seed=41)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    scales = np.geomspace(0.5, 2.0, 4)
    got = lmm.logpdf_sweep(f(lmm.MOInputIsotopicByOutputs(x, p), 0.1), y, scales)
    for s, v in zip(scales, got):
        fs_s = [o.GP(o.Kernel(g.kernel.kind, g.kernel.variance, g.kernel.inv_lengthscale * s), g.mean_const) for g in fs]
        assert rel(v, o.oilmm_logpdf(o.OILMMModel(fs_s, U, S), x, 0.1, y)) < RTOL


def test_errors(lmm):
    x, xs, U, S, fs, y = make_problem(20, 3, 2, 2, seed=1)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    with pytest.raises(RuntimeError, match="out dim of x != out dim of f."):
        lmm.logpdf(f(lmm.MOInputIsotopicByOutputs(x, 4), 0.1), np.zeros(80))
    with pytest.raises(ValueError, match="not an orthogonal matrix"):
        lmm.Orthogonal(np.random.default_rng(0).uniform(size=(3, 2)), S)
    with pytest.raises(TypeError):
        lmm.logpdf(f(lmm.MOInputIsotopicByFeatures(x, 3), 0.1), y)  # MethodError in Julia: src/ilmm.jl:45
    # duplicate inputs and vanishing noise: singular latent covariance -> PosDefException
    xd = np.array([0.0, 0.0, 1.0])
    fi = lmm.independent_mogp([lmm.GP(lmm.SEKernel())])
    with pytest.raises(lmm.PosDefException):
        lmm.logpdf(fi(lmm.MOInputIsotopicByOutputs(xd, 1), 1e-300), np.zeros(3))


def test_device_resident_inputs(lmm):
    """x / y may already live in HBM (bench.py's `value` arm): same result as host inputs."""
    import torch

    x, xs, U, S, fs, y = make_problem(500, 4, 2, 3, seed=51)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    a = lmm.logpdf(f(lmm.MOInputIsotopicByOutputs(x, 4), 0.1), y)
    xd = torch.from_numpy(x.reshape(-1, 1)).cuda()
    yd = torch.from_numpy(y).cuda()
    b = lmm.logpdf(f(lmm.MOInputIsotopicByOutputs(xd, 4), 0.1), yd)
    assert a == b


def test_cov_and_mean_and_cov(lmm):
    """`cov` / `mean_and_cov` (src/ilmm.jl:132-139,147; src/independent_mogp.jl:60-63): prior and
    posterior, OILMM / general ILMM / IndependentMOGP -- the quantities test_utils.jl:50-59 compares."""
    rng = np.random.default_rng(12)
    N, Ns, p, m = 60, 9, 3, 2
    x = np.sort(rng.uniform(0, 4, N))
    xs = np.sort(rng.uniform(0, 4, Ns))
    U, S = o.orthogonal_from_seed(p, m, seed=4)
    fs = [o.GP(o.Kernel(o.SE), 0.3), o.GP(o.Kernel(o.MATERN32, 0.7, 1.3), -0.1)]
    om = o.OILMMModel(fs, U, S)
    H = om.H
    y = rng.standard_normal(p * N)
    gps = [to_lmm_gp(lmm, g) for g in fs]
    xin, xsin = lmm.MOInputIsotopicByOutputs(x, p), lmm.MOInputIsotopicByOutputs(xs, p)
    f_o = lmm.ILMM(lmm.independent_mogp(gps), lmm.Orthogonal(U, S))
    f_i = lmm.ILMM(lmm.independent_mogp(gps), H)
    # prior
    Mr, Cr = o.ilmm_mean_and_cov(fs, H, xs, 0.1)
    for f in (f_o, f_i):
        M, C = lmm.mean_and_cov(f(xsin, 0.1))
        assert_isapprox(M, Mr, RTOL)
        np.testing.assert_allclose(C, Cr, rtol=RTOL, atol=1e-12)
        np.testing.assert_allclose(np.diag(lmm.cov(f(xsin, 0.1))), lmm.var(f(xsin, 0.1)), rtol=1e-12)
    # OILMM posterior
    post = lmm.posterior(f_o(xin, 0.1), y)
    opost = o.oilmm_posterior(om, x, 0.1, y)
    Mr, Cr = o.ilmm_mean_and_cov(opost.fs, H, xs, 0.1)
    M, C = lmm.mean_and_cov(post(xsin, 0.1))
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(C, Cr, rtol=1e-8, atol=1e-11)
    np.testing.assert_allclose(np.diag(C), lmm.var(post(xsin, 0.1)), rtol=1e-9)
    # general ILMM posterior (joint)
    posti = lmm.posterior(f_i(xin, 0.1), y)
    Mr, Cr = o.ilmm_mean_and_cov(o.ilmm_posterior(fs, H, x, 0.1, y), H, xs, 0.1)
    M, C = lmm.mean_and_cov(posti(xsin, 0.1))
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(C, Cr, rtol=1e-8, atol=1e-11)
    # IndependentMOGP prior / posterior, by outputs and by features
    fm = lmm.independent_mogp(gps)
    xs2 = lmm.MOInputIsotopicByOutputs(xs, 2)
    Mr, Cr = o.imogp_mean_and_cov(fs, xs, 0.2)
    M, C = lmm.mean_and_cov(fm(xs2, 0.2))
    np.testing.assert_allclose(M, Mr, rtol=RTOL)
    np.testing.assert_allclose(C, Cr, rtol=RTOL, atol=1e-13)
    y2 = rng.standard_normal(2 * N)
    pm = lmm.posterior(fm(lmm.MOInputIsotopicByOutputs(x, 2), 0.1), y2)
    Mr, Cr = o.imogp_mean_and_cov(o.imogp_posterior(fs, x, 0.1, y2), xs, 0.2)
    M, C = lmm.mean_and_cov(pm(xs2, 0.2))
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(C, Cr, rtol=1e-8, atol=1e-11)
    idx = o.indices_outputs_to_features(Ns, 2)
    Mf, Cf = lmm.mean_and_cov(pm(lmm.MOInputIsotopicByFeatures(xs, 2), 0.2))
    np.testing.assert_allclose(Cf, Cr[np.ix_(idx, idx)], rtol=1e-8, atol=1e-11)


def test_ilmm_posterior_rand_and_logpdf(lmm):
    """`rand(rng, p_i)` and `logpdf(p_i(x*, σ²), y*)` on a general-ILMM posterior (src/ilmm.jl:78-87,
    150-163 with PosteriorGP{IndependentMOGP} latents; test/ilmm.jl:26-27, notebook `rand(rng, p_i)`)."""
    rng = np.random.default_rng(14)
    N, Ns, p, m = 40, 7, 3, 2
    x = np.sort(rng.uniform(0, 6, N))
    xs = np.sort(rng.uniform(0, 6, Ns))
    H = rng.uniform(0, 1, (p, m))
    fs = [o.GP(o.Kernel(o.MATERN32), 0.2), o.GP(o.Kernel(o.MATERN52, 0.8, 1.4), -0.3)]
    y = rng.standard_normal(p * N)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), H)
    post = lmm.posterior(f(lmm.MOInputIsotopicByOutputs(x, p), 0.1), y)
    opost = o.ilmm_posterior(fs, H, x, 0.1, y)
    pfx = post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1)
    g = np.random.default_rng(31)
    zl, zn = g.standard_normal(m * Ns), g.standard_normal(p * Ns)
    s = lmm.rand(np.random.default_rng(31), pfx)
    np.testing.assert_allclose(s, o.ilmm_post_rand(opost, xs, 0.1, zl, zn), rtol=1e-6, atol=1e-7)
    ys = rng.standard_normal(p * Ns)
    assert rel(lmm.logpdf(pfx, ys), o.ilmm_post_logpdf(opost, xs, 0.1, ys)) < RTOL


def test_sequential_conditioning(lmm):
    """posterior(post(x2, σ²), y2): OILMM and IndependentMOGP posteriors conditioned a
    second (and third) time agree with the textbook update of the first posterior, and with
    conditioning once on all the data."""
    rng = np.random.default_rng(17)
    N1, N2, N3, Nt, p, m = 90, 37, 5, 11, 4, 2
    x1, x2, x3 = np.sort(rng.uniform(0, 6, N1)), rng.uniform(0, 6, N2), rng.uniform(0, 6, N3)
    xt = rng.uniform(0, 6, Nt)
    U, S = o.orthogonal_from_seed(p, m, seed=6)
    fs = [o.GP(o.Kernel(o.SE, 1.0, 1.2), 0.2), o.GP(o.Kernel(o.MATERN52, 0.6, 0.9), -0.4)]
    om = o.OILMMModel(fs, U, S)
    y1, y2, y3 = rng.standard_normal(p * N1), rng.standard_normal(p * N2), rng.standard_normal(p * N3)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    mo = lambda x: lmm.MOInputIsotopicByOutputs(x, p)
    post1 = lmm.posterior(f(mo(x1), 0.1), y1)
    post2 = lmm.posterior(post1(mo(x2), 0.3), y2)
    assert isinstance(post2, lmm.OILMM)
    M, V = lmm.mean_and_var(post2(mo(xt), 0.1))
    # oracle: latent-wise textbook update of the first posterior, then the OILMM mixing
    op1 = o.oilmm_posterior(om, x1, 0.1, y1)
    T, ST2 = o.project_orthogonal(U, S, 0.3)
    Ty2 = T @ o.reshape_y(y2, N2)
    ML, VL = zip(*[o.gp_condition_again_marginals(op1.fs[i], x2, ST2[i], Ty2[i], xt) for i in range(m)])
    H = om.H
    Mr = (H @ np.stack(ML)).reshape(-1)
    Vr = ((H * H) @ (np.stack(VL) + 1e-18) + 0.1).reshape(-1)
    np.testing.assert_allclose(M, Mr, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(V, Vr, rtol=1e-8)
    # same noise both times == conditioning once on the union (by-outputs concatenation per output)
    post2b = lmm.posterior(post1(mo(x2), 0.1), y2)
    xu = np.concatenate([x1, x2])
    yu = np.concatenate([np.concatenate([y1.reshape(p, N1)[j], y2.reshape(p, N2)[j]]) for j in range(p)])
    postu = lmm.posterior(f(mo(xu), 0.1), yu)
    Ma, Va = lmm.mean_and_var(post2b(mo(xt), 0.1))
    Mb, Vb = lmm.mean_and_var(postu(mo(xt), 0.1))
    np.testing.assert_allclose(Ma, Mb, rtol=1e-9, atol=1e-11)
    np.testing.assert_allclose(Va, Vb, rtol=1e-9)
    # a third conditioning step keeps the per-point noise of the first two
    post3 = lmm.posterior(post2(mo(x3), 0.05), y3)
    M3, V3 = lmm.mean_and_var(post3(mo(xt), 0.1))
    assert M3.shape == (p * Nt,) and np.all(V3 > 0.1) and np.all(V3 <= V + 1e-12)
    # IndependentMOGP
    fm = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    ya, yb = rng.standard_normal(m * N1), rng.standard_normal(m * N2)
    pa = lmm.posterior(fm(lmm.MOInputIsotopicByOutputs(x1, m), 0.1), ya)
    pb = lmm.posterior(pa(lmm.MOInputIsotopicByOutputs(x2, m), 0.2), yb)
    Mi, Vi = lmm.mean_and_var(pb(lmm.MOInputIsotopicByOutputs(xt, m), 0.1))
    oa = o.imogp_posterior(fs, x1, 0.1, ya)
    for i in range(m):
        mr, vr = o.gp_condition_again_marginals(oa[i], x2, 0.2, yb.reshape(m, N2)[i], xt)
        np.testing.assert_allclose(Mi[i * Nt:(i + 1) * Nt], mr, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(Vi[i * Nt:(i + 1) * Nt], vr + 0.1, rtol=1e-8)


@pytest.mark.parametrize("N,p,m", [(60, 4, 3), (700, 6, 3), (1300, 4, 2)])
def test_logpdf_gradient(lmm, N, p, m):
    """rrule of logpdf: value + gradients w.r.t. kernel hyper-parameters, σ² and y from
    the batched potri + fused kernel-gradient reduction, against the oracle's analytic gradient."""
    x, xs, U, S, fs, y = make_problem(N, p, m, 1, seed=61 + N, means=True)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    lp, g = lmm.logpdf_and_gradient(f(lmm.MOInputIsotopicByOutputs(x, p), 0.15), y, with_grad_y=True)
    lpr, gr = o.oilmm_logpdf_grad(om, x, 0.15, y)
    assert rel(lp, lpr) < RTOL
    for k in ("variance", "inv_lengthscale", "mean_const"):
        np.testing.assert_allclose(g[k], gr[k], rtol=1e-7, atol=1e-8)
    assert rel(g["sigma2"], gr["sigma2"]) < 1e-7
    np.testing.assert_allclose(g["y"], gr["y"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(g["S"], gr["S"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(g["U"], gr["U"], rtol=1e-7, atol=1e-7)
    # IndependentMOGP
    fi = lmm.independent_mogp([to_lmm_gp(lmm, gg) for gg in fs])
    yi = np.random.default_rng(3).standard_normal(m * N)
    lpi, gi = lmm.logpdf_and_gradient(fi(lmm.MOInputIsotopicByOutputs(x, m), 0.15), yi, with_grad_y=True)
    parts = [o.gp_logpdf_grad(fs[i], x, 0.15, yi.reshape(m, N)[i]) for i in range(m)]
    assert rel(lpi, sum(q[0] for q in parts)) < RTOL
    np.testing.assert_allclose(gi["variance"], [q[1] for q in parts], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(gi["inv_lengthscale"], [q[2] for q in parts], rtol=1e-7, atol=1e-8)
    assert rel(gi["sigma2"], sum(q[4] for q in parts)) < 1e-7
    np.testing.assert_allclose(gi["y"], np.concatenate([q[5] for q in parts]), rtol=1e-7, atol=1e-9)


@pytest.mark.parametrize("N,p,m,D", [(40, 4, 3, 1), (150, 5, 2, 2), (300, 6, 3, 1)])
def test_ilmm_logpdf_gradient(lmm, N, p, m, D):
    """rrule of the general-ILMM logpdf (test/ilmm.jl:31 `gradient(logpdf, ilmmx, y_train)`): joint potri +
    block-wise kernel-gradient contraction + host chain through `project`, against the oracle's analytic
    gradient (itself checked against finite differences on CPU)."""
    x, xs, U, S, fs, y = make_problem(N, p, m, 1, seed=71 + N, D=D, means=True)
    H = np.random.default_rng(N).uniform(0.0, 1.0, (p, m))
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), H)
    xin = x if D == 1 else lmm.ColVecs(x.T)
    lp, g = lmm.logpdf_and_gradient(f(lmm.MOInputIsotopicByOutputs(xin, p), 0.2), y, with_grad_y=True)
    lpr, gr = o.ilmm_logpdf_grad(fs, H, x, 0.2, y)
    assert rel(lp, lpr) < RTOL
    for k in ("variance", "inv_lengthscale", "mean_const"):
        np.testing.assert_allclose(g[k], gr[k], rtol=1e-7, atol=1e-8)
    assert rel(g["sigma2"], gr["sigma2"]) < 1e-7
    np.testing.assert_allclose(g["y"], gr["y"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(g["H"], gr["H"], rtol=1e-7, atol=1e-7)


@pytest.mark.parametrize("N,Ns,p,m", [(50, 7, 4, 3), (300, 140, 5, 2)])
def test_posterior_logpdf_gradient(lmm, N, Ns, p, m):
    """`gradient(logpdf, po, y_test)` (test/oilmm.jl:32, test/ilmm.jl:32, test/independent_mogp.jl:66): value and
    gradient w.r.t. σ² and y* of the posterior-predictive logpdf for all three model kinds."""
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=83 + N, means=True)
    rng = np.random.default_rng(N)
    ys = rng.standard_normal(p * Ns)
    lat = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    O = lmm.MOInputIsotopicByOutputs
    # OILMM
    po = lmm.posterior(lmm.ILMM(lat, lmm.Orthogonal(U, S))(O(x, p), 0.1), y)
    lp, g = lmm.logpdf_and_gradient(po(O(xs, p), 0.2), ys, with_grad_y=True)
    lpr, gr = o.oilmm_post_logpdf_grad(o.oilmm_posterior(o.OILMMModel(fs, U, S), x, 0.1, y), xs, 0.2, ys)
    assert rel(lp, lpr) < RTOL
    assert rel(g["sigma2"], gr["sigma2"]) < 1e-7
    np.testing.assert_allclose(g["y"], gr["y"], rtol=1e-7, atol=1e-9)
    # IndependentMOGP
    pi = lmm.posterior(lat(O(x, m), 0.1), y[: m * N])
    lp, g = lmm.logpdf_and_gradient(pi(O(xs, m), 0.2), ys[: m * Ns], with_grad_y=True)
    lpr, gr = o.oilmm_post_logpdf_grad(o.OILMMModel(o.imogp_posterior(fs, x, 0.1, y[: m * N]), np.eye(m), np.ones(m)), xs, 0.2, ys[: m * Ns])
    assert rel(lp, lpr) < RTOL
    assert rel(g["sigma2"], gr["sigma2"]) < 1e-7
    np.testing.assert_allclose(g["y"], gr["y"], rtol=1e-7, atol=1e-9)
    # general ILMM
    H = rng.uniform(0.0, 1.0, (p, m))
    pl = lmm.posterior(lmm.ILMM(lat, H)(O(x, p), 0.1), y)
    lp, g = lmm.logpdf_and_gradient(pl(O(xs, p), 0.2), ys, with_grad_y=True)
    lpr, gr = o.ilmm_post_logpdf_grad(o.ilmm_posterior(fs, H, x, 0.1, y), xs, 0.2, ys)
    assert rel(lp, lpr) < 1e-8
    assert rel(g["sigma2"], gr["sigma2"]) < 1e-6
    np.testing.assert_allclose(g["y"], gr["y"], rtol=1e-6, atol=1e-8)


@pytest.mark.parametrize("N1,N2,Nt,p,m", [(12, 9, 5, 4, 3), (150, 140, 33, 5, 2)])
def test_ilmm_sequential_conditioning(lmm, N1, N2, Nt, p, m):
    """posterior(post(x2, σ2²), y2) on a general-ILMM posterior (src/ilmm.jl:184-198 with PosteriorGP{IndependentMOGP}
    latents) against the textbook update of the joint latent posterior, and against the posterior on the union."""
    rng = np.random.default_rng(N1)
    x1, x2, xt = np.sort(rng.uniform(0, 4, N1)), rng.uniform(0, 4, N2), rng.uniform(0, 4, Nt)
    _, _, _, _, fs, _ = make_problem(N1, p, m, 1, seed=5, means=True)
    H = rng.uniform(0, 1, (p, m))
    y1, y2 = rng.standard_normal(p * N1), rng.standard_normal(p * N2)
    O = lmm.MOInputIsotopicByOutputs
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), H)
    p1 = lmm.posterior(f(O(x1, p), 0.1), y1)
    p2, lp2 = lmm.posterior(p1(O(x2, p), 0.2), y2, with_logpdf=True)
    M, V = lmm.mean_and_var(p2(O(xt, p), 0.15))
    op = o.ilmm_posterior(fs, H, x1, 0.1, y1)
    Mr, Vr = o.ilmm_condition_again_mean_and_var(op, x2, 0.2, y2, xt, 0.15)
    np.testing.assert_allclose(M, Mr, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(V, Vr, rtol=1e-8)
    assert rel(lp2, o.ilmm_post_logpdf(op, x2, 0.2, y2)) < 1e-8
    # a third conditioning step keeps working (per-point noise blocks are extended again)
    p3 = lmm.posterior(p2(O(xt, p), 0.3), rng.standard_normal(p * Nt))
    M3, V3 = lmm.mean_and_var(p3(O(xt, p), 0.1))
    assert np.all(np.isfinite(M3)) and np.all(V3 > 0.1) and np.all(V3 < V + 1.0)


@pytest.mark.parametrize("N,Ns,m", [(9, 4, 2), (140, 37, 3)])
def test_imogp_vector_and_dense_noise(lmm, N, Ns, m):
    """f(x_mo, v::Vector) and f(x_mo, Σy::Matrix) on an IndependentMOGP: the AbstractGPs generic FiniteGP path
    (test/independent_mogp.jl:72-75; by-features reordering of a Diagonal Σy src/independent_mogp.jl:149-159)."""
    rng = np.random.default_rng(7 * N)
    x, xs = np.sort(rng.uniform(0, 4, N)), rng.uniform(0, 4, Ns)
    _, _, _, _, fs, _ = make_problem(N, m, m, 1, seed=9, means=True)
    f = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    y, ys = rng.standard_normal(m * N), rng.standard_normal(m * Ns)
    O, F = lmm.MOInputIsotopicByOutputs, lmm.MOInputIsotopicByFeatures
    v = rng.uniform(0.05, 0.5, m * N)
    A = rng.standard_normal((m * N, m * N))
    Sy = A.T @ A / (m * N) + 0.1 * np.eye(m * N)
    for Sig in (v, Sy):
        fx = f(O(x, m), Sig)
        assert rel(lmm.logpdf(fx, y), o.imogp_logpdf_noise(fs, x, Sig, y)) < RTOL
        post, lp = lmm.posterior(fx, y, with_logpdf=True)
        assert rel(lp, o.imogp_logpdf_noise(fs, x, Sig, y)) < RTOL
        Mr, Cr = o.imogp_posterior_noise_mean_and_cov(fs, x, Sig, y, xs, 0.2)
        M, V = lmm.mean_and_var(post(O(xs, m), 0.2))
        np.testing.assert_allclose(M, Mr, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(V, np.diag(Cr), rtol=1e-8)
        M2, C2 = lmm.mean_and_cov(post(O(xs, m), 0.2))
        np.testing.assert_allclose(C2, Cr, rtol=1e-7, atol=1e-9)
        # logpdf(post(x*, σ²), y*) == MVN(y*; mean*, cov* + σ² I)
        L = np.linalg.cholesky(Cr)
        z = sla.solve_triangular(L, ys - Mr, lower=True)
        ref = -0.5 * (m * Ns * o.LOG2PI + 2 * np.sum(np.log(np.diag(L))) + z @ z)
        assert rel(lmm.logpdf(post(O(xs, m), 0.2), ys), ref) < 1e-8
        assert lmm.rand(np.random.default_rng(0), post(O(xs, m), 0.2)).shape == (m * Ns,)
        assert lmm.rand(np.random.default_rng(0), fx).shape == (m * N,)
        # prior marginals under the same noise: var = k(x,x) + diag(Σy)
        Mp, Vp = lmm.mean_and_var(fx)
        np.testing.assert_allclose(Vp, np.concatenate([np.full(N, g.kernel.variance) for g in fs]) + (Sig if Sig.ndim == 1 else np.diag(Sig)), rtol=1e-13)
    # by-features inputs: Σy given in by-features order
    idx = o.indices_outputs_to_features(N, m)
    assert rel(lmm.logpdf(f(F(x, m), v[idx]), y[idx]), o.imogp_logpdf_noise(fs, x, v, y)) < RTOL
    assert rel(lmm.logpdf(f(F(x, m), Sy[np.ix_(idx, idx)]), y[idx]), o.imogp_logpdf_noise(fs, x, Sy, y)) < RTOL
    pf = lmm.posterior(f(F(x, m), v[idx]), y[idx])
    Mf, Vf = lmm.mean_and_var(pf(F(xs, m), 0.2))
    Mr, Cr = o.imogp_posterior_noise_mean_and_cov(fs, x, v, y, xs, 0.2)
    ids = o.indices_outputs_to_features(Ns, m)
    np.testing.assert_allclose(Mf, Mr[ids], rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(Vf, np.diag(Cr)[ids], rtol=1e-8)
    # ILMM / OILMM methods keep rejecting non-scalar noise (src/ilmm.jl:45 dispatch)
    with pytest.raises(TypeError):
        lmm.ILMM(f, np.eye(m))(O(x, m), v)


def test_posterior_save_load_round_trip(lmm, tmp_path):
    """Serialisable posterior: a handle restored from its file answers bit-identically, for the
    per-latent kinds (OILMM, IndependentMOGP), the joint kind (general ILMM) and a sequentially conditioned one."""
    N, Ns, p, m = 260, 40, 5, 3
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=101, means=True)
    O = lmm.MOInputIsotopicByOutputs
    lat = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    H = np.random.default_rng(2).uniform(0, 1, (p, m))
    models = [(lmm.ILMM(lat, lmm.Orthogonal(U, S)), p), (lat, m), (lmm.ILMM(lat, H), p)]
    for k, (model, pp) in enumerate(models):
        post = lmm.posterior(model(O(x, pp), 0.1), y[: pp * N])
        if k != 1:
            post = lmm.posterior(post(O(xs, pp), 0.2), y[: pp * Ns])  # also a conditioned-again handle
        path = str(tmp_path / f"post{k}.lmm")
        lmm.save_posterior(post, path)
        back = lmm.load_posterior(path, model)
        M0, V0 = lmm.mean_and_var(post(O(xs, pp), 0.1))
        M1, V1 = lmm.mean_and_var(back(O(xs, pp), 0.1))
        assert np.array_equal(M0, M1) and np.array_equal(V0, V1)
        assert lmm.logpdf(post(O(xs, pp), 0.1), y[: pp * Ns]) == lmm.logpdf(back(O(xs, pp), 0.1), y[: pp * Ns])
        again = lmm.posterior(back(O(xs, pp), 0.3), y[: pp * Ns])  # a loaded handle can be conditioned further
        assert np.all(np.isfinite(lmm.mean(again(O(xs, pp), 0.1))))
    with open(str(tmp_path / "bad.lmm"), "wb") as fh:
        fh.write(b"not a posterior")
    with pytest.raises(Exception):
        lmm.load_posterior(str(tmp_path / "bad.lmm"), lat)


def test_ard_transform(lmm, tmp_path):
    """`k ∘ ARDTransform(v)` (KernelFunctions: x -> v .* x before distances) on D = 3 inputs through every
    kernel-evaluating path: OILMM logpdf / posterior / marginals / gradient, general ILMM (joint assembly + cross covariance),
    sequential conditioning and a save / load round trip (the handle keeps its own copy of the ARD vectors)."""
    N, Ns, p, m, D = 300, 37, 5, 3, 3
    rng = np.random.default_rng(17)
    x, xs = rng.uniform(0, 3, (N, D)), rng.uniform(0, 3, (Ns, D))
    U, S = o.orthogonal_from_seed(p, m, seed=3)
    fs = [o.GP(o.Kernel(o.SE, 0.9, 1.2, (0.5, 1.5, 1.0)), 0.3), o.GP(o.Kernel(o.MATERN32, 1.3, 0.8, (1.2, 0.7, 2.0)), -0.2),
          o.GP(o.Kernel(o.MATERN52, 0.7, 1.1))]  # the third latent has no ARD
    y, ys = rng.standard_normal(p * N), rng.standard_normal(p * Ns)
    O = lmm.MOInputIsotopicByOutputs
    lat = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lat, lmm.Orthogonal(U, S))
    fx = f(O(lmm.RowVecs(x), p), 0.1)
    post, lp = lmm.posterior(fx, y, with_logpdf=True)
    assert rel(lp, o.oilmm_logpdf(om, x, 0.1, y)) < RTOL
    opost = o.oilmm_posterior(om, x, 0.1, y)
    M, V = lmm.mean_and_var(post(O(lmm.RowVecs(xs), p), 0.1))
    Mr, Vr = o.oilmm_mean_and_var(opost, xs, 0.1)
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(V, Vr, rtol=RTOL)
    lpg, g = lmm.logpdf_and_gradient(fx, y, with_grad_y=True)
    lpr, gr = o.oilmm_logpdf_grad(om, x, 0.1, y)
    for k in ("variance", "inv_lengthscale", "mean_const"):
        np.testing.assert_allclose(g[k], gr[k], rtol=1e-7, atol=1e-8)
    assert rel(g["sigma2"], gr["sigma2"]) < 1e-7
    T, ST1 = o.project_orthogonal(U, S, 0.1)
    ard_ref = np.stack([o.gp_logpdf_grad_ard(fs[i], x, ST1[i], (T @ y.reshape(p, N))[i]) for i in range(m)])
    np.testing.assert_allclose(g["ard"], ard_ref, rtol=1e-7, atol=1e-8)
    assert np.all(g["ard"][2] == 0.0)  # the latent without an ARDTransform
    # save / load keeps the ARD vectors; the reloaded handle can be conditioned again
    path = str(tmp_path / "ard.lmm")
    lmm.save_posterior(post, path)
    back = lmm.load_posterior(path, f)
    M1, V1 = lmm.mean_and_var(back(O(lmm.RowVecs(xs), p), 0.1))
    assert np.array_equal(M, M1) and np.array_equal(V, V1)
    post2 = lmm.posterior(back(O(lmm.RowVecs(xs), p), 0.2), ys)
    # compare with the textbook per-latent update of the first posterior
    T, ST2 = o.project_orthogonal(U, S, 0.2)
    xt = rng.uniform(0, 3, (9, D))
    M2, V2 = lmm.mean_and_var(post2(O(lmm.RowVecs(xt), p), 0.1))
    ML, VL = [], []
    for i in range(m):
        mr, vr = o.gp_condition_again_marginals(opost.fs[i], xs, ST2[i], (T @ ys.reshape(p, Ns))[i], xt)
        ML.append(mr)
        VL.append(vr + 1e-18)
    H = om.H
    np.testing.assert_allclose(M2, (H @ np.stack(ML)).reshape(-1), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(V2, ((H * H) @ np.stack(VL) + 0.1).reshape(-1), rtol=1e-8)
    # general ILMM: joint assembly and block-diagonal cross covariance evaluate the same kernels
    Hd = rng.uniform(0, 1, (p, m))
    fi = lmm.ILMM(lat, Hd)(O(lmm.RowVecs(x), p), 0.1)
    assert rel(lmm.logpdf(fi, y), o.ilmm_logpdf(fs, Hd, x, 0.1, y)) < RTOL
    pi = lmm.posterior(fi, y)
    Mi, Vi = lmm.mean_and_var(pi(O(lmm.RowVecs(xs), p), 0.1))
    Mir, Vir = o.ilmm_mean_and_var(o.ilmm_posterior(fs, Hd, x, 0.1, y), Hd, xs, 0.1)
    np.testing.assert_allclose(Mi, Mir, rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(Vi, Vir, rtol=1e-8)
    lpi, gi = lmm.logpdf_and_gradient(fi, y)
    _, gir = o.ilmm_logpdf_grad(fs, Hd, x, 0.1, y)
    np.testing.assert_allclose(gi["inv_lengthscale"], gir["inv_lengthscale"], rtol=1e-7, atol=1e-8)
    np.testing.assert_allclose(gi["ard"], gir["ard"], rtol=1e-7, atol=1e-8)
    # D > 8 with ARD is rejected loudly, never routed elsewhere
    big = lmm.GP(lmm.SEKernel().compose(lmm.ARDTransform(np.ones(9))))
    with pytest.raises(ValueError, match="ARDTransform"):
        lmm.logpdf(lmm.independent_mogp([big])(O(lmm.RowVecs(rng.uniform(0, 1, (5, 9))), 1), 0.1), np.zeros(5))


@pytest.mark.parametrize("D", [1, 2])
def test_exponential_and_rational_quadratic_kernels(lmm, D):
    """ExponentialKernel (= Matern12Kernel) and RationalQuadraticKernel(α) latents: OILMM logpdf, posterior marginals and the
    logpdf gradient, and the general-ILMM joint assembly, against the oracle."""
    N, Ns, p, m = 260, 40, 4, 3
    rng = np.random.default_rng(23 + D)
    x = np.sort(rng.uniform(0, 4, N)) if D == 1 else rng.uniform(0, 3, (N, D))
    xs = rng.uniform(0, 4, Ns) if D == 1 else rng.uniform(0, 3, (Ns, D))
    U, S = o.orthogonal_from_seed(p, m, seed=5)
    ard = None if D == 1 else (0.8, 1.3)
    fs = [o.GP(o.Kernel(o.EXPONENTIAL, 0.9, 1.2, ard), 0.3), o.GP(o.Kernel(o.RATQUAD, 1.3, 0.8, None, 1.7), -0.2),
          o.GP(o.Kernel(o.RATQUAD, 0.7, 1.1, ard, 0.6))]
    y = rng.standard_normal(p * N)
    O = lmm.MOInputIsotopicByOutputs
    wrap = (lambda a: a) if D == 1 else lmm.RowVecs
    lat = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    om = o.OILMMModel(fs, U, S)
    fx = lmm.ILMM(lat, lmm.Orthogonal(U, S))(O(wrap(x), p), 0.1)
    post, lp = lmm.posterior(fx, y, with_logpdf=True)
    assert rel(lp, o.oilmm_logpdf(om, x, 0.1, y)) < RTOL
    M, V = lmm.mean_and_var(post(O(wrap(xs), p), 0.1))
    Mr, Vr = o.oilmm_mean_and_var(o.oilmm_posterior(om, x, 0.1, y), xs, 0.1)
    assert_isapprox(M, Mr, RTOL)
    np.testing.assert_allclose(V, Vr, rtol=RTOL)
    _, g = lmm.logpdf_and_gradient(fx, y)
    _, gr = o.oilmm_logpdf_grad(om, x, 0.1, y)
    for k in ("variance", "inv_lengthscale", "mean_const"):
        np.testing.assert_allclose(g[k], gr[k], rtol=1e-7, atol=1e-8)
    Hd = rng.uniform(0, 1, (p, m))
    assert rel(lmm.logpdf(lmm.ILMM(lat, Hd)(O(wrap(x), p), 0.1), y), o.ilmm_logpdf(fs, Hd, x, 0.1, y)) < RTOL
    with pytest.raises(ValueError):
        lmm.RationalQuadraticKernel(0.0)


@pytest.mark.parametrize("N,Ns,p,m,frac", [(30, 6, 3, 2, 0.3), (300, 45, 5, 3, 0.4), (260, 33, 4, 4, 0.0)])
def test_missing_data_ilmm(lmm, N, Ns, p, m, frac):
    """Heterotopic / missing-data conditioning: NaN entries of y are unobserved; logpdf of the observed entries and
    posterior marginals of all outputs against textbook conditioning of the dense multi-output GP; with nothing missing the
    result equals the ILMM's own dense form."""
    rng = np.random.default_rng(N + 1)
    x, xs = np.sort(rng.uniform(0, 5, N)), rng.uniform(0, 5, Ns)
    _, _, U, S, fs, y = make_problem(N, p, m, 1, seed=7 + N, means=True)
    ym = y.copy()
    ym[rng.uniform(size=p * N) < frac] = np.nan
    if frac > 0:
        ym[: N // 2] = np.nan  # half of output 1 missing in one block
    O = lmm.MOInputIsotopicByOutputs
    lat = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    for Hobj, Hd in ((np.random.default_rng(3).uniform(0, 1, (p, m)), None), (lmm.Orthogonal(U, S), U * np.sqrt(S)[None, :])):
        Hd = np.asarray(Hobj) if Hd is None else Hd
        fx = lmm.ILMM(lat, Hobj)(O(x, p), 0.1)
        post, lp = lmm.posterior_missing(fx, ym, with_logpdf=True)
        if isinstance(post, lmm.OILMM):  # Orthogonal H and a per-input mask (here: nothing missing): the structured OILMM path
            assert frac == 0.0 and isinstance(Hobj, lmm.Orthogonal)
        else:
            assert post.n_observed == int(np.sum(~np.isnan(ym)))
        assert rel(lp, o.missing_data_logpdf(fs, Hd, x, 0.1, ym)) < RTOL
        assert rel(lmm.logpdf_missing(fx, ym), lp) < 1e-14
        M, V = lmm.mean_and_var(post(O(xs, p), 0.2))
        Mr, Vr = o.missing_data_posterior_mean_and_var(fs, Hd, x, 0.1, ym, xs, 0.2)
        np.testing.assert_allclose(M, Mr, rtol=1e-8, atol=1e-10)
        np.testing.assert_allclose(V, Vr, rtol=1e-8)
        assert len(lmm.marginals(post(O(xs, p), 0.2))) == p * Ns
    if frac == 0.0:  # nothing missing: the dense form of the ILMM itself
        assert rel(lp, o.dense_mogp_logpdf(fs, Hd, x, 0.1, y)) < RTOL
    with pytest.raises(ValueError):
        lmm.posterior_missing(fx, np.full(p * N, np.nan))


def test_imogp_process_cov_mixed_orderings(lmm):
    """cov(f, x, y) with by-outputs / by-features inputs in all four combinations
    (src/independent_mogp.jl:60-71,181-215; test/independent_mogp.jl:135-141)."""
    rng = np.random.default_rng(19)
    xa, xb = rng.uniform(0, 3, 5), rng.uniform(0, 3, 140)
    fs = [o.GP(o.Kernel(o.SE, 0.5)), o.GP(o.Kernel(o.MATERN32, 1.0, 1.4)), o.GP(o.Kernel(o.MATERN52))]
    f = lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs])
    import scipy.linalg as sla2
    ref = sla2.block_diag(*[o.kernelmatrix(g.kernel, xa, xb) for g in fs])
    ia, ib = o.indices_outputs_to_features(5, 3), o.indices_outputs_to_features(140, 3)
    O, F = lmm.MOInputIsotopicByOutputs, lmm.MOInputIsotopicByFeatures
    np.testing.assert_allclose(lmm.cov(f, O(xa, 3), O(xb, 3)), ref, rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(lmm.cov(f, F(xa, 3), O(xb, 3)), ref[ia, :], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(lmm.cov(f, O(xa, 3), F(xb, 3)), ref[:, ib], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(lmm.cov(f, F(xa, 3), F(xb, 3)), ref[np.ix_(ia, ib)], rtol=1e-13, atol=1e-15)
    np.testing.assert_allclose(lmm.cov(f, F(xa, 3)), sla2.block_diag(*[o.kernelmatrix(g.kernel, xa) for g in fs])[np.ix_(ia, ia)], rtol=1e-13, atol=1e-15)


@pytest.mark.parametrize("N,m", [(1, 2), (129, 3), (700, 2), (1300, 5), (2600, 3)])
def test_persistent_solve_sweeps_match_stepwise_and_oracle(lmm, N, m):
    """The persistent forward / backward sweep kernels (one launch per direction, tile rows chained through ready flags)
    against the one-launch-per-tile-column kernels and the oracle: quadratic form (through the lml terms), α and the
    posterior mean.  Includes a one-tile case, ragged sizes and 21 tile columns."""
    p, Ns = m + 1, 33
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=N + 7, means=True)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    ctx = lmm.default_context()
    res = {}
    try:
        for impl in (0, 1):
            ctx.set_option("solve_impl", impl)
            post, lp = lmm.posterior(fx, y, with_logpdf=True)
            res[impl] = (lp, post.f.fs[m - 1].alpha, lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))[0])
    finally:
        ctx.set_option("solve_impl", 1)
    assert rel(res[1][0], res[0][0]) < 1e-13
    assert_isapprox(res[1][1], res[0][1], 1e-11, "alpha: sweep vs stepwise")
    opost = o.oilmm_posterior(om, x, 0.1, y)
    assert rel(res[1][0], o.oilmm_logpdf(om, x, 0.1, y)) < RTOL
    assert_isapprox(res[1][1], opost.fs[m - 1].alpha, RTOL, "alpha vs oracle")
    assert_isapprox(res[1][2], o.oilmm_mean_and_var(opost, xs, 0.1)[0], RTOL, "posterior mean vs oracle")


@pytest.mark.parametrize("N1,N2", [(700, 300), (768, 40), (1300, 1), (200, 1100)])
def test_block_cholesky_update_conditioning(lmm, N1, N2):
    """lmm_post_condition extends each latent's factor by a block-Cholesky update (AbstractGPs' sequential conditioning,
    test/oilmm.jl:20-26) instead of re-factorising the union: same factor / α as the union path, and the
    marginals match the oracle's textbook update of the first posterior at 1e-9.  N1 below, at and above a tile boundary."""
    import time

    rng = np.random.default_rng(N1 + N2)
    p, m, Nt = 4, 3, 29
    x1, x2, xt = np.sort(rng.uniform(0, 9, N1)), rng.uniform(0, 9, N2), rng.uniform(0, 9, Nt)
    U, S = o.orthogonal_from_seed(p, m, seed=6)
    fs = [o.GP(o.Kernel(o.SE, 1.0, 1.2), 0.2), o.GP(o.Kernel(o.MATERN52, 0.6, 0.9), -0.4), o.GP(o.Kernel(o.MATERN32, 1.3, 0.7), 0.0)]
    om = o.OILMMModel(fs, U, S)
    y1, y2 = rng.standard_normal(p * N1), rng.standard_normal(p * N2)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    mo = lambda x: lmm.MOInputIsotopicByOutputs(x, p)
    ctx = lmm.default_context()
    post1 = lmm.posterior(f(mo(x1), 0.1), y1)
    res = {}
    try:
        for upd in (1, 0):
            ctx.set_option("condition_update", upd)
            t0 = time.perf_counter()
            post2 = lmm.posterior(post1(mo(x2), 0.3), y2)
            dt = time.perf_counter() - t0
            g = post2.f.fs[m - 1]
            res[upd] = (g.C, g.alpha, lmm.mean_and_var(post2(mo(xt), 0.1)), dt)
    finally:
        ctx.set_option("condition_update", 1)
    assert_isapprox(res[1][0], res[0][0], 1e-12, "extended factor vs union factor")
    assert_isapprox(res[1][1], res[0][1], 1e-11, "alpha: update vs union")
    # the rows of the old factor are carried over bit for bit
    jrows = (N1 // 128) * 128
    assert np.array_equal(res[1][0][:jrows, :jrows], post1.f.fs[m - 1].C[:jrows, :jrows])
    op1 = o.oilmm_posterior(om, x1, 0.1, y1)
    T, ST2 = o.project_orthogonal(U, S, 0.3)
    Ty2 = T @ o.reshape_y(y2, N2)
    ML, VL = zip(*[o.gp_condition_again_marginals(op1.fs[i], x2, ST2[i], Ty2[i], xt) for i in range(m)])
    M, V = res[1][2]
    assert_isapprox(M, (om.H @ np.stack(ML)).reshape(-1), RTOL, "mean after block update vs oracle")
    np.testing.assert_allclose(V, ((om.H * om.H) @ (np.stack(VL) + 1e-18) + 0.1).reshape(-1), rtol=1e-8)


def test_block_update_is_cheap(lmm):
    """Conditioning a C3-sized posterior (N = 8192, here 4 latents) on 128 more points costs a small fraction of a fresh
    factorisation (< 5 % at the full m = 16, measured by tools/bench_configs.py; asserted here: < 25 %)."""
    import time

    rng = np.random.default_rng(1)
    N1, N2, p, m = 8192, 128, 8, 4
    x1, x2 = np.sort(rng.uniform(0, 80, N1)), rng.uniform(0, 80, N2)
    U, S = o.orthogonal_from_seed(p, m, seed=6)
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.Matern52Kernel().compose(lmm.ScaleTransform(1.0 + 0.2 * i))) for i in range(m)]), lmm.Orthogonal(U, S))
    mo = lambda x: lmm.MOInputIsotopicByOutputs(x, p)
    y1, y2 = rng.standard_normal(p * N1), rng.standard_normal(p * N2)
    ctx = lmm.default_context()
    import os

    ctx.set_option("ozaki", 0)  # like against like: the block update runs on DMMA, so the fresh factorisation it is compared with does too
    post1 = lmm.posterior(f(mo(x1), 0.1), y1)
    lmm.posterior(f(mo(x1), 0.1), y1)
    fresh_ms = float(ctx.last_timings()[0])
    ctx.set_option("ozaki", int(os.environ.get("LMM_OZAKI", "0")))
    best = 1e30
    for _ in range(3):
        t0 = time.perf_counter()
        post2 = lmm.posterior(post1(mo(x2), 0.1), y2)
        best = min(best, (time.perf_counter() - t0) * 1e3)
    assert best < 0.25 * fresh_ms, (best, fresh_ms)


@pytest.mark.parametrize("N,p,m", [(50, 3, 2), (552, 600, 20), (1000, 64, 64), (300, 37, 5), (2050, 130, 33)])
def test_project_dmma_matches_scalar_kernel_and_oracle(lmm, N, p, m):
    """The FP64 tensor-core projection kernel (T*Y, (I - UU')Y, residual norm; every column-block width and the p-row split
    used at small N) against the round-1 scalar-FMA kernel and the oracle: projected rows δ_i, regulariser, logpdf.
    (552, 600, 20) is the reference's published notebook shape."""
    x, xs, U, S, fs, y = make_problem(N, p, m, 5, seed=N + p, means=True)
    om = o.OILMMModel(fs, U, S)
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    ctx = lmm.default_context()
    res = {}
    try:
        for impl in (0, 1):
            ctx.set_option("project_impl", impl)
            post = lmm.posterior(fx, y)
            res[impl] = (lmm.logpdf_terms(fx, y), post.f.fs[0].delta, post.f.fs[m - 1].delta)
    finally:
        ctx.set_option("project_impl", 1)
    T, ST = o.project_orthogonal(U, S, 0.1)
    Y = o.reshape_y(y, N)
    assert_isapprox(res[1][1], T[0] @ Y - fs[0].mean_const, 1e-12, "projected row 0")
    assert_isapprox(res[1][2], T[m - 1] @ Y - fs[m - 1].mean_const, 1e-12, "projected last row")
    assert_isapprox(res[1][1], res[0][1], 1e-12, "DMMA vs scalar projection")
    reg = o.regulariser_orthogonal(U, S, 0.1, Y)
    assert rel(res[1][0][m], reg) < RTOL and rel(res[0][0][m], reg) < RTOL
    ref_terms, _ = o.oilmm_logpdf_terms(om, x, 0.1, y)
    np.testing.assert_allclose(res[1][0][:m], ref_terms, rtol=RTOL)


def test_missing_data_cov_rand_save_load(lmm, tmp_path):
    """The missing-data (dense-model) posterior beyond marginals: mean_and_cov / cov against textbook
    conditioning of the dense multi-output GP, rand = mean + chol(C + σ²I) z for a caller-supplied z, and a save / load round
    trip that answers bit-identically."""
    rng = np.random.default_rng(21)
    N, Ns, p, m = 70, 9, 4, 2
    x, xs = np.sort(rng.uniform(0, 7, N)), rng.uniform(0, 7, Ns)
    H = rng.uniform(0, 1, (p, m))
    fs = [o.GP(o.Kernel(o.SE, 1.1, 0.9), 0.3), o.GP(o.Kernel(o.MATERN32, 0.8, 1.2), -0.1)]
    y = rng.standard_normal(p * N)
    ym = y.copy()
    ym[rng.uniform(size=p * N) < 0.3] = np.nan
    O = lmm.MOInputIsotopicByOutputs
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), H)
    post = lmm.posterior_missing(f(O(x, p), 0.1), ym)
    M, Cm = lmm.mean_and_cov(post(O(xs, p), 0.2))
    Mr, Cr = o.missing_data_posterior_mean_and_cov(fs, H, x, 0.1, ym, xs, 0.2)
    assert_isapprox(M, Mr, RTOL, "missing-data posterior mean")
    assert_isapprox(Cm, Cr, 1e-9, "missing-data posterior covariance")
    assert np.array_equal(Cm, Cm.T)
    np.testing.assert_allclose(np.diag(Cm), lmm.var(post(O(xs, p), 0.2)), rtol=1e-10)

    class FixedNormals:  # a Generator stand-in that hands rand a known z
        def __init__(self, z):
            self.z = z

        def standard_normal(self, n):
            assert n == len(self.z)
            return self.z

    z = rng.standard_normal(p * Ns)
    import lmm_b200.api as api

    s = api._rand_one(FixedNormals(z), post(O(xs, p), 0.2))
    assert_isapprox(s, Mr + np.linalg.cholesky(Cr) @ z, 1e-8, "missing-data posterior sample")
    path = str(tmp_path / "post_masked.lmm")
    lmm.save_posterior(post, path)
    back = lmm.load_posterior(path, f)
    Mb, Vb = lmm.mean_and_var(back(O(xs, p), 0.2))
    M0, V0 = lmm.mean_and_var(post(O(xs, p), 0.2))
    assert np.array_equal(Mb, M0) and np.array_equal(Vb, V0)


def test_heterotopic_oilmm_with_whole_inputs_missing(lmm):
    """An OILMM whose mask is per input (whole time steps missing) stays an OILMM on the observed inputs: the structured path
    (per-latent factors, full posterior API) equals the dense missing-data model and the oracle, and a mask that is not
    per-input falls back to the dense model."""
    rng = np.random.default_rng(8)
    N, Ns, p, m = 150, 12, 5, 3
    x, xs, U, S, fs, y = make_problem(N, p, m, Ns, seed=31, means=True)
    gone = rng.uniform(size=N) < 0.35
    ym = y.reshape(p, N).copy()
    ym[:, gone] = np.nan
    ym = ym.reshape(-1)
    O = lmm.MOInputIsotopicByOutputs
    f = lmm.ILMM(lmm.independent_mogp([to_lmm_gp(lmm, g) for g in fs]), lmm.Orthogonal(U, S))
    post, lp = lmm.posterior_missing(f(O(x, p), 0.1), ym, with_logpdf=True)
    assert isinstance(post, lmm.OILMM)  # structured: an ordinary OILMM posterior over the observed inputs
    om = o.OILMMModel(fs, U, S)
    xo, yo = x[~gone], y.reshape(p, N)[:, ~gone].reshape(-1)
    assert rel(lp, o.oilmm_logpdf(om, xo, 0.1, yo)) < RTOL
    Hd = om.H
    assert rel(lp, o.missing_data_logpdf(fs, Hd, x, 0.1, ym)) < 1e-8  # == the dense model on the observed entries
    assert rel(lmm.logpdf_missing(f(O(x, p), 0.1), ym), lp) < 1e-14
    M, V = lmm.mean_and_var(post(O(xs, p), 0.1))
    Mr, Vr = o.oilmm_mean_and_var(o.oilmm_posterior(om, xo, 0.1, yo), xs, 0.1)
    assert_isapprox(M, Mr, RTOL, "heterotopic OILMM posterior mean")
    np.testing.assert_allclose(V, Vr, rtol=RTOL)
    Md, Vd = o.missing_data_posterior_mean_and_var(fs, Hd, x, 0.1, ym, xs, 0.1)
    assert_isapprox(M, Md, 1e-7, "structured vs dense missing-data mean")
    # the full API on the structured posterior: rand and sequential conditioning
    assert lmm.rand(np.random.default_rng(0), post(O(xs, p), 0.1)).shape == (p * Ns,)
    post2 = lmm.posterior(post(O(xs, p), 0.1), rng.standard_normal(p * Ns))
    assert isinstance(post2, lmm.OILMM)
    # one output missing at one input: not per-input any more -> dense model
    ym2 = ym.copy()
    ym2[int(np.flatnonzero(~gone)[0])] = np.nan
    post_d = lmm.posterior_missing(f(O(x, p), 0.1), ym2)
    assert not isinstance(post_d, lmm.OILMM)
    Mdd, _ = lmm.mean_and_var(post_d(O(xs, p), 0.1))
    Mrr, _ = o.missing_data_posterior_mean_and_var(fs, Hd, x, 0.1, ym2, xs, 0.1)
    assert_isapprox(Mdd, Mrr, 1e-8, "dense fallback mean")
