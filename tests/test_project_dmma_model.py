"""CPU check of the index maps of project_dmma_kernel (csrc/proj.cu) through its lane-level NumPy model
(tools/project_dmma_model.py): fragment addresses, task split, epilogue rows / columns, p-row split, ragged p / m / N."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))
import project_dmma_model as pm  # noqa: E402


@pytest.mark.parametrize("p,m,N,NB,psplit,lat0,mloc", [(3, 2, 50, 8, 1, 0, 2), (13, 5, 37, 16, 2, 0, 5), (24, 20, 70, 32, 3, 4, 9),
                                                       (40, 9, 130, 64, 1, 0, 9), (64, 64, 96, 64, 4, 32, 32), (17, 17, 9, 8, 4, 0, 17)])
def test_model_matches_dense(p, m, N, NB, psplit, lat0, mloc):
    rng = np.random.default_rng(p * 1000 + N)
    Y = rng.standard_normal((p, N))
    U, _ = np.linalg.qr(rng.standard_normal((p, m)))
    S = rng.uniform(0.5, 2.0, m)
    T = U.T / np.sqrt(S)[:, None]
    means = rng.standard_normal(mloc)
    ty, ss, z, R = pm.project(Y, T, lat0, mloc, means, U.T.copy(), U, NB, psplit)
    np.testing.assert_allclose(ty, (T @ Y)[lat0:lat0 + mloc] - means[:, None], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(z, U.T @ Y, rtol=1e-12, atol=1e-13)
    Rref = Y - U @ (U.T @ Y)
    np.testing.assert_allclose(R, Rref, rtol=1e-10, atol=1e-12)
    assert ss == pytest.approx(float(np.sum(Rref * Rref)), rel=1e-10, abs=1e-20)


def test_model_general_ilmm_shares_one_product():
    """P is T (general ILMM): Z = T Y is both the projection and the operand of the residual Y - H (T Y)."""
    rng = np.random.default_rng(5)
    p, m, N = 11, 4, 45
    Y = rng.standard_normal((p, N))
    H = rng.uniform(0, 1, (p, m))
    T = np.linalg.solve(H.T @ H / 0.1 + 1e-9 * np.eye(m), H.T / 0.1)
    means = rng.standard_normal(m)
    ty, ss, z, R = pm.project(Y, T, 0, m, means, T, H, 16, 2)
    np.testing.assert_allclose(ty, T @ Y - means[:, None], rtol=1e-12, atol=1e-13)
    np.testing.assert_allclose(R, Y - H @ (T @ Y), rtol=1e-10, atol=1e-12)
