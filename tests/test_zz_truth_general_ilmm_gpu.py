"""The CUDA path against the second 40-digit fixture (tests/golden/c2_truth_mp.npz, make_golden_mp2.py): a general ILMM with a
dense non-orthogonal H, 2-D inputs, an ARDTransform, Matern52 / Exponential / RationalQuadratic latents and constant means.
liblmm follows the reference's projected algorithm (src/ilmm.jl:61-68: T = (H'H/σ² + 1e-9 I)⁻¹ H'/σ²), so it inherits the
reference's own 1e-9 regulariser offset from the exact model -- the tolerances below are the ones the CPU oracle's restatement
of the same algorithm meets against the same truth (tests/test_golden.py)."""
import os

import numpy as np
import pytest

T2 = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c2_truth_mp.npz"))


@pytest.mark.gpu
def test_cuda_general_ilmm_matches_extended_precision_truth():
    import lmm_b200 as lmm

    x, xs, H, y, s2 = T2["x"], T2["xs"], T2["H"], T2["y"], float(T2["sigma2"])
    p = H.shape[0]
    k0 = (1.2 * lmm.Matern52Kernel()).compose(lmm.ScaleTransform(0.9)).compose(lmm.ARDTransform(T2["ard0"]))
    k1 = (0.8 * lmm.ExponentialKernel()).compose(lmm.ScaleTransform(1.4))
    k2 = lmm.RationalQuadraticKernel(1.7).compose(lmm.ScaleTransform(0.6))
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP(0.5, k0), lmm.GP(-1.0, k1), lmm.GP(0.0, k2)]), H)
    O = lmm.MOInputIsotopicByOutputs
    fx = f(O(lmm.RowVecs(x), p), s2)
    lp = lmm.logpdf(fx, y)
    assert abs(lp - float(T2["logpdf"])) <= 1e-9 * abs(float(T2["logpdf"]))
    M, V = lmm.mean_and_var(lmm.posterior(fx, y)(O(lmm.RowVecs(xs), p), s2))
    np.testing.assert_allclose(M, T2["post_mean"], rtol=1e-7, atol=1e-9)
    np.testing.assert_allclose(V, T2["post_var"], rtol=1e-8)
    # against the oracle's restatement of the SAME (projected, 1e-9-jittered) algorithm the north star's 1e-9 holds
    from oracle import lmm_oracle as o
    from _tol import assert_isapprox

    fs = [o.GP(o.Kernel(o.MATERN52, 1.2, 0.9, ard=tuple(T2["ard0"])), 0.5), o.GP(o.Kernel(o.EXPONENTIAL, 0.8, 1.4), -1.0),
          o.GP(o.Kernel(o.RATQUAD, 1.0, 0.6, param=1.7), 0.0)]
    assert abs(lp - o.ilmm_logpdf(fs, H, x, s2, y)) <= 1e-9 * abs(lp)
    Mo, Vo = o.ilmm_mean_and_var(o.ilmm_posterior(fs, H, x, s2, y), H, xs, s2)
    assert_isapprox(M, Mo, 1e-9, "general-ILMM posterior mean vs projected oracle")
    assert_isapprox(V, Vo, 1e-9, "general-ILMM posterior var vs projected oracle")
    _, g = lmm.logpdf_and_gradient(fx, y)
    assert abs(g["sigma2"] - float(T2["dlogpdf_dsigma2"])) <= 1e-7 * abs(float(T2["dlogpdf_dsigma2"]))
    assert abs(g["ard"][0][1] - float(T2["dlogpdf_dard0_1"])) <= 1e-6 * abs(float(T2["dlogpdf_dard0_1"]))
    assert abs(g["H"][1, 2] - float(T2["dlogpdf_dH12"])) <= 1e-6 * abs(float(T2["dlogpdf_dH12"]))
