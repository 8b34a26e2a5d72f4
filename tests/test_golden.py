"""Golden vectors (tests/golden/c1_oilmm.npz, made by tests/golden/make_golden.py from the oracle
with an extended-precision cross-check): the oracle must keep reproducing them (CPU), and the CUDA
path must match them (GPU)."""
import os

import numpy as np
import pytest

from _tol import assert_isapprox

from oracle import lmm_oracle as o

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_oilmm.npz"))
FS = [o.GP(o.Kernel(o.SE)), o.GP(o.Kernel(o.MATERN32))]
FS_I = [o.GP(o.Kernel(o.MATERN32), 30.0), o.GP(o.Kernel(o.SE, 0.5), 10.0)]


def test_oracle_reproduces_golden():
    model = o.OILMMModel(FS, G["U"], G["S"])
    terms, reg = o.oilmm_logpdf_terms(model, G["x"], 0.1, G["y"])
    np.testing.assert_allclose(terms, G["lml_terms"], rtol=1e-13)
    assert float(np.sum(terms) + reg) == pytest.approx(float(G["logpdf"]), rel=1e-13)
    assert float(G["logpdf"]) == pytest.approx(float(G["logpdf_longdouble"]), rel=1e-11)
    post = o.oilmm_posterior(model, G["x"], 0.1, G["y"])
    M, V = o.oilmm_mean_and_var(post, G["xs"], 0.1)
    np.testing.assert_allclose(M, G["post_mean"], rtol=1e-11, atol=1e-13)
    np.testing.assert_allclose(V, G["post_var"], rtol=1e-11)
    assert o.oilmm_logpdf(post, G["xs"], 0.1, G["ys"]) == pytest.approx(float(G["post_logpdf"]), rel=1e-12)
    assert o.imogp_logpdf(FS_I, G["x"], 0.1, G["imogp_y"]) == pytest.approx(float(G["imogp_logpdf"]), rel=1e-13)
    assert o.ilmm_logpdf(FS, G["ilmm_H"], G["x"], 0.1, G["y"]) == pytest.approx(float(G["ilmm_logpdf"]), rel=1e-12)


@pytest.mark.gpu
def test_cuda_path_matches_golden():
    import lmm_b200 as lmm

    p = 3
    mo = lambda x, q=p: lmm.MOInputIsotopicByOutputs(x, q)
    gps = [lmm.GP(lmm.SEKernel()), lmm.GP(lmm.Matern32Kernel())]
    f = lmm.ILMM(lmm.independent_mogp(gps), lmm.Orthogonal(G["U"], G["S"]))
    fx = f(mo(G["x"]), 0.1)
    terms = lmm.logpdf_terms(fx, G["y"])
    np.testing.assert_allclose(terms[:2], G["lml_terms"], rtol=1e-9)
    assert abs(lmm.logpdf(fx, G["y"]) - float(G["logpdf"])) <= 1e-9 * abs(float(G["logpdf"]))
    post = lmm.posterior(fx, G["y"])
    M, V = lmm.mean_and_var(post(mo(G["xs"]), 0.1))
    assert_isapprox(M, G["post_mean"], 1e-9)
    np.testing.assert_allclose(V, G["post_var"], rtol=1e-9)
    got = lmm.logpdf(post(mo(G["xs"]), 0.1), G["ys"])
    assert abs(got - float(G["post_logpdf"])) <= 1e-9 * abs(float(G["post_logpdf"]))
    fi = lmm.independent_mogp([lmm.GP(30.0, lmm.Matern32Kernel()), lmm.GP(10.0, 0.5 * lmm.SEKernel())])
    fxi = fi(mo(G["x"], 2), 0.1)
    assert abs(lmm.logpdf(fxi, G["imogp_y"]) - float(G["imogp_logpdf"])) <= 1e-9 * abs(float(G["imogp_logpdf"]))
    Mi, Vi = lmm.mean_and_var(lmm.posterior(fxi, G["imogp_y"])(mo(G["xs"], 2), 0.1))
    np.testing.assert_allclose(Mi, G["imogp_post_mean"], rtol=1e-9)
    np.testing.assert_allclose(Vi, G["imogp_post_var"], rtol=1e-9)
    fg = lmm.ILMM(lmm.independent_mogp(gps), G["ilmm_H"])
    fxg = fg(mo(G["x"]), 0.1)
    assert abs(lmm.logpdf(fxg, G["y"]) - float(G["ilmm_logpdf"])) <= 1e-9 * abs(float(G["ilmm_logpdf"]))
    Mg, Vg = lmm.mean_and_var(lmm.posterior(fxg, G["y"])(mo(G["xs"]), 0.1))
    assert_isapprox(Mg, G["ilmm_post_mean"], 1e-9)
    np.testing.assert_allclose(Vg, G["ilmm_post_var"], rtol=1e-9)


# ---- 40-digit truth (tests/golden/c1_truth_mp.npz, made by make_golden_mp.py with mpmath straight from the dense multi-output
# GP definition the reference's tests use as ground truth; shares no code with the oracle)
T = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c1_truth_mp.npz"))


def test_oracle_matches_extended_precision_truth():
    model = o.OILMMModel(FS, G["U"], G["S"])
    assert o.oilmm_logpdf(model, G["x"], 0.1, G["y"]) == pytest.approx(float(T["logpdf"]), rel=1e-13)
    M, V = o.oilmm_mean_and_var(o.oilmm_posterior(model, G["x"], 0.1, G["y"]), G["xs"], 0.1)
    np.testing.assert_allclose(M, T["post_mean"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(V, T["post_var"], rtol=1e-11)
    _, g = o.oilmm_logpdf_grad(model, G["x"], 0.1, G["y"])
    assert g["sigma2"] == pytest.approx(float(T["dlogpdf_dsigma2"]), rel=1e-10)
    assert g["inv_lengthscale"][0] == pytest.approx(float(T["dlogpdf_dinv_lengthscale0"]), rel=1e-10)
    assert g["variance"][0] == pytest.approx(float(T["dlogpdf_dvariance0"]), rel=1e-10)


@pytest.mark.gpu
def test_cuda_path_matches_extended_precision_truth():
    """The B200 path against the 40-digit values: logpdf, posterior means and variances to 1e-9 relative (the north star's
    tolerance), the logpdf gradient to 1e-8."""
    import lmm_b200 as lmm

    mo = lambda x: lmm.MOInputIsotopicByOutputs(x, 3)
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.SEKernel()), lmm.GP(lmm.Matern32Kernel())]), lmm.Orthogonal(G["U"], G["S"]))
    fx = f(mo(G["x"]), 0.1)
    lp = lmm.logpdf(fx, G["y"])
    assert abs(lp - float(T["logpdf"])) <= 1e-9 * abs(float(T["logpdf"]))
    M, V = lmm.mean_and_var(lmm.posterior(fx, G["y"])(mo(G["xs"]), 0.1))
    assert_isapprox(M, T["post_mean"], 1e-9)
    np.testing.assert_allclose(V, T["post_var"], rtol=1e-9)
    _, g = lmm.logpdf_and_gradient(fx, G["y"])
    assert abs(g["sigma2"] - float(T["dlogpdf_dsigma2"])) <= 1e-8 * abs(float(T["dlogpdf_dsigma2"]))
    assert abs(g["inv_lengthscale"][0] - float(T["dlogpdf_dinv_lengthscale0"])) <= 1e-8 * abs(float(T["dlogpdf_dinv_lengthscale0"]))
    assert abs(g["variance"][0] - float(T["dlogpdf_dvariance0"])) <= 1e-8 * abs(float(T["dlogpdf_dvariance0"]))


# ---- second 40-digit truth (tests/golden/c2_truth_mp.npz, make_golden_mp2.py): a general ILMM with a dense non-orthogonal H,
# 2-D inputs, an ARDTransform, Matern52 / Exponential / RationalQuadratic latents and constant means
T2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_truth_mp.npz"))


def _c2_latents():
    return [o.GP(o.Kernel(o.MATERN52, 1.2, 0.9, ard=tuple(T2["ard0"])), 0.5), o.GP(o.Kernel(o.EXPONENTIAL, 0.8, 1.4), -1.0),
            o.GP(o.Kernel(o.RATQUAD, 1.0, 0.6, param=1.7), 0.0)]


def test_oracle_general_ilmm_matches_extended_precision_truth():
    """The oracle's dense multi-output GP (what test/ilmm.jl compares the ILMM with) reproduces the 40-digit values to
    rounding; its restatement of the reference's projected ILMM (src/ilmm.jl:61-68,150-198) agrees to the size of the
    reference's own 1e-9 regulariser in `project` -- that offset belongs to the reference's algorithm, not to the oracle."""
    fs, x, xs, H, y, s2 = _c2_latents(), T2["x"], T2["xs"], T2["H"], T2["y"], float(T2["sigma2"])
    assert o.dense_mogp_logpdf(fs, H, x, s2, y) == pytest.approx(float(T2["logpdf"]), rel=1e-13)
    M, V = o.dense_mogp_posterior_mean_and_var(fs, H, x, s2, y, xs, s2)
    np.testing.assert_allclose(M, T2["post_mean"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(V, T2["post_var"], rtol=1e-11)
    assert o.ilmm_logpdf(fs, H, x, s2, y) == pytest.approx(float(T2["logpdf"]), rel=1e-9)
    M2, V2 = o.ilmm_mean_and_var(o.ilmm_posterior(fs, H, x, s2, y), H, xs, s2)
    np.testing.assert_allclose(M2, T2["post_mean"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(V2, T2["post_var"], rtol=1e-8)
    _, g = o.ilmm_logpdf_grad(fs, H, x, s2, y)
    assert g["sigma2"] == pytest.approx(float(T2["dlogpdf_dsigma2"]), rel=1e-8)
    assert g["ard"][0][1] == pytest.approx(float(T2["dlogpdf_dard0_1"]), rel=1e-7)
    assert g["H"][1, 2] == pytest.approx(float(T2["dlogpdf_dH12"]), rel=1e-7)
