"""Host-side mirror of the reference's types (linearmixingmodels.jl_b200/api.py): everything that can be checked without a
GPU -- kernel algebra (ScaledKernel, ScaleTransform, ARDTransform composition as KernelFunctions defines them), FiniteGP
noise handling, input wrappers, dispatch errors (the reference's MethodErrors), by-features index permutations."""
import numpy as np
import pytest

import lmm_b200 as lmm
from oracle import lmm_oracle as o


def test_kernel_algebra_matches_kernelfunctions_semantics():
    k = 0.5 * lmm.SEKernel()
    assert (k.kind, k.variance, k.inv_lengthscale, k.ard) == (0, 0.5, 1.0, None)
    k2 = (2.0 * k).compose(lmm.ScaleTransform(3.0))  # scaling multiplies, transforms compose multiplicatively
    assert k2.variance == 1.0 and k2.inv_lengthscale == 3.0
    assert lmm.with_lengthscale(lmm.Matern52Kernel(), 4.0).inv_lengthscale == 0.25
    ka = lmm.Matern32Kernel().compose(lmm.ARDTransform([1.0, 2.0])).compose(lmm.ARDTransform([3.0, 0.5]))
    assert ka.ard == (3.0, 1.0)
    with pytest.raises(ValueError):
        ka.compose(lmm.ARDTransform([1.0, 2.0, 3.0]))
    rq = 1.5 * lmm.RationalQuadraticKernel(0.7).compose(lmm.ScaleTransform(2.0))
    assert (rq.kind, rq.param, rq.variance, rq.inv_lengthscale) == (4, 0.7, 1.5, 2.0)
    assert lmm.Matern12Kernel().kind == lmm.ExponentialKernel().kind == 3
    with pytest.raises(TypeError):
        -1.0 * lmm.SEKernel()
    # the descriptor table handed to the C ABI
    d = lmm.api._descs([lmm.GP(3.0, rq), lmm.GP(ka)])
    assert (d[0].kind, d[0].variance, d[0].inv_lengthscale, d[0].mean_const, d[0].param, d[0].ard) == (4, 1.5, 2.0, 3.0, 0.7, None)
    assert d[1].ard is not None and len(d._keep) == 1 and tuple(d._keep[0]) == (3.0, 1.0)


def test_gp_and_model_constructors():
    g = lmm.GP(lmm.SEKernel())
    assert g.mean_const == 0.0 and lmm.GP(10, lmm.SEKernel()).mean_const == 10.0
    fs = lmm.independent_mogp([g, lmm.GP(lmm.Matern32Kernel())])
    assert isinstance(fs, lmm.IndependentMOGP) and len(fs.fs) == 2
    U, S = o.orthogonal_from_seed(3, 2, seed=1)
    f = lmm.ILMM(fs, lmm.Orthogonal(U, S))
    assert isinstance(f, lmm.OILMM) and not isinstance(lmm.ILMM(fs, np.ones((3, 2))), lmm.OILMM)  # OILMM is a dispatch alias (src/oilmm.jl:13)
    assert lmm.get_latent_gp(f) is fs
    assert np.asarray(lmm.Orthogonal(U, np.diag(S))).shape == (3, 2)  # Diagonal(S) given as a matrix


def test_finitegp_noise_forms_and_dispatch_errors():
    fs = lmm.independent_mogp([lmm.GP(lmm.SEKernel()), lmm.GP(lmm.SEKernel())])
    x = lmm.MOInputIsotopicByOutputs(np.linspace(0, 1, 4), 2)
    assert len(x) == 8
    assert fs(x).sigma2 == 1e-18  # AbstractGPs default FiniteGP noise
    fv = fs(x, np.full(8, 0.1))
    assert fv.noise.shape == (8,) and fv.sigma2 == 0.0
    fm = fs(x, np.eye(8))
    assert fm.noise.shape == (8, 8)
    with pytest.raises(ValueError):
        fs(x, np.ones(7))
    with pytest.raises(TypeError):  # ILMM / OILMM methods dispatch on scalar noise only (src/ilmm.jl:45)
        lmm.ILMM(fs, np.eye(2))(x, np.full(8, 0.1))
    with pytest.raises(TypeError):  # and on by-outputs inputs only
        lmm.unpack(lmm.ILMM(fs, np.eye(2))(lmm.MOInputIsotopicByFeatures(np.linspace(0, 1, 4), 2), 0.1))
    # the noise is carried to by-outputs order for a by-features FiniteGP (src/independent_mogp.jl:149-151)
    xf = lmm.MOInputIsotopicByFeatures(np.linspace(0, 1, 4), 2)
    v = np.arange(8.0)
    kind, vo = lmm.api._noise_by_outputs(fs(xf, v))
    assert kind == 1 and np.array_equal(vo, v[o.indices_features_to_outputs(4, 2)])
    A = np.arange(64.0).reshape(8, 8)
    kind, Ao = lmm.api._noise_by_outputs(fs(xf, A))
    idx = o.indices_features_to_outputs(4, 2)
    assert kind == 2 and np.array_equal(Ao, A[np.ix_(idx, idx)])


def test_input_wrappers():
    X = np.arange(6.0).reshape(2, 3)  # D = 2, N = 3
    assert lmm.ColVecs(X).points().shape == (3, 2) and np.array_equal(lmm.ColVecs(X).points()[1], X[:, 1])
    assert np.array_equal(lmm.RowVecs(X.T).points(), lmm.ColVecs(X).points())
    assert np.array_equal(lmm.reshape_y(np.arange(6.0), 3), np.arange(6.0).reshape(2, 3))


def test_reorder_indices_round_trip():
    x = lmm.MOInputIsotopicByFeatures(np.linspace(0, 1, 5), 3)
    a, b = lmm.indices_which_reorder_outputs_to_features(x), lmm.indices_which_reorder_features_to_outputs(x)
    v = np.arange(15)
    assert np.array_equal(v[a][b], v) and np.array_equal(v[b][a], v)
    assert np.array_equal(a, o.indices_outputs_to_features(5, 3)) and np.array_equal(b, o.indices_features_to_outputs(5, 3))
