"""Tolerance helpers shared by the parity tests.

BASELINE.json's north star asks for 1e-9 *relative* on logpdf, posterior means and posterior variances.  For vectors
that is Julia's `isapprox(a, b; rtol)`: norm(a - b) <= rtol * max(norm(a), norm(b)) -- no absolute floor, so
near-zero entries of a posterior mean cannot hide behind an `atol`."""
import numpy as np


def relnorm(a, b) -> float:
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    den = max(np.linalg.norm(a), np.linalg.norm(b))
    return float(np.linalg.norm(a - b) / den) if den > 0 else 0.0


def assert_isapprox(a, b, rtol=1e-9, what=""):
    r = relnorm(a, b)
    assert r <= rtol, f"{what} norm-relative error {r:.3e} > {rtol:.1e}"
