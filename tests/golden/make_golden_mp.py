#!/usr/bin/env python
"""Generates tests/golden/c1_truth_mp.npz: 40-digit (mpmath) evaluations of BASELINE config 1 written directly from the
definitions -- the dense (pN x pN) multi-output GP  y ~ N((H ⊗ I) m, Σ_i (h_i h_i') ⊗ K_i + σ² I)  that the reference's own
tests use as the ground truth of the OILMM (test/oilmm.jl:10-14 via test/ilmm.jl:5 `LinearMixingModelKernel`) -- sharing no
code with oracle/lmm_oracle.py beyond the input generator: logpdf, posterior marginals at x*, and the derivatives of the logpdf
w.r.t. σ², the first latent's inverse lengthscale and variance by 40-digit central differences.  These are "truth to 1e-30":
when the CUDA path and the NumPy oracle disagree in the last digits, this says which one is closer.

    python tests/golden/make_golden_mp.py        (about a minute)
"""
import os
import sys

import mpmath as mp
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lmm_oracle as o  # noqa: E402  (only orthogonal_from_seed: the inputs)

mp.mp.dps = 40


def kern(kind, d):
    if kind == "se":
        return mp.exp(-d * d / 2)
    s = mp.sqrt(3) * d  # matern32
    return (1 + s) * mp.exp(-s)


def cov_dense(xa, xb, H, params, noise=None):
    """Σ_i H[:,i] H[:,i]' ⊗ K_i(xa, xb), by outputs; params[i] = (kind, variance, inv_lengthscale)."""
    p, m = len(H), len(H[0])
    na, nb = len(xa), len(xb)
    Ks = []
    for (kind, var, s) in params:
        Ks.append([[var * kern(kind, abs(s * (a - b))) for b in xb] for a in xa])
    C = mp.zeros(p * na, p * nb)
    for j in range(p):
        for j2 in range(p):
            for i in range(m):
                w = H[j][i] * H[j2][i]
                for a in range(na):
                    for b in range(nb):
                        C[j * na + a, j2 * nb + b] += w * Ks[i][a][b]
    if noise is not None:
        for k in range(p * na):
            C[k, k] += noise
    return C


def chol(C):
    n = C.rows
    L = mp.zeros(n, n)
    for j in range(n):
        d = C[j, j] - mp.fsum(L[j, k] ** 2 for k in range(j))
        L[j, j] = mp.sqrt(d)
        for i in range(j + 1, n):
            L[i, j] = (C[i, j] - mp.fsum(L[i, k] * L[j, k] for k in range(j))) / L[j, j]
    return L


def fwd(L, b):
    n = L.rows
    z = [mp.mpf(0)] * n
    for i in range(n):
        z[i] = (b[i] - mp.fsum(L[i, k] * z[k] for k in range(i))) / L[i, i]
    return z


def bwd(L, b):
    n = L.rows
    z = [mp.mpf(0)] * n
    for i in reversed(range(n)):
        z[i] = (b[i] - mp.fsum(L[k, i] * z[k] for k in range(i + 1, n))) / L[i, i]
    return z


def logpdf(x, y, H, params, sigma2):
    L = chol(cov_dense(x, x, H, params, sigma2))
    z = fwd(L, y)
    n = len(y)
    return -(n * mp.log(2 * mp.pi) + 2 * mp.fsum(mp.log(L[i, i]) for i in range(n)) + mp.fsum(v * v for v in z)) / 2, L


def main():
    rng = np.random.default_rng(20240416)  # the same inputs as make_golden.py (config 1, zero-mean latents)
    N, p, m, Ns = 50, 3, 2, 7
    x = np.sort(rng.uniform(0, 5, N))
    xs = rng.uniform(0, 5, Ns)
    U, S = o.orthogonal_from_seed(p, m, seed=1)
    y = rng.standard_normal(p * N)
    mpf = lambda a: mp.mpf(float(a))  # the Float64 inputs, exactly
    H = [[mpf(U[j, i]) * mp.sqrt(mpf(S[i])) for i in range(m)] for j in range(p)]
    xm, xsm, ym = [mpf(v) for v in x], [mpf(v) for v in xs], [mpf(v) for v in y]
    params = [("se", mp.mpf(1), mp.mpf(1)), ("m32", mp.mpf(1), mp.mpf(1))]
    s2 = mpf(0.1)
    lp, L = logpdf(xm, ym, H, params, s2)
    alpha = bwd(L, fwd(L, ym))
    Ksx = cov_dense(xsm, xm, H, params)
    Kss = cov_dense(xsm, xsm, H, params)
    mean, var = [], []
    for r in range(p * Ns):
        row = [Ksx[r, c] for c in range(p * N)]
        mean.append(mp.fsum(a * b for a, b in zip(row, alpha)))
        v = fwd(L, row)
        var.append(Kss[r, r] - mp.fsum(t * t for t in v) + s2)  # predictive noise σ² = 0.1 (src/oilmm.jl:72)
    h = mp.mpf(10) ** -12
    d_s2 = (logpdf(xm, ym, H, params, s2 + h)[0] - logpdf(xm, ym, H, params, s2 - h)[0]) / (2 * h)
    pp = lambda dv, ds: [("se", 1 + dv, 1 + ds), params[1]]
    d_s = (logpdf(xm, ym, H, pp(0, h), s2)[0] - logpdf(xm, ym, H, pp(0, -h), s2)[0]) / (2 * h)
    d_v = (logpdf(xm, ym, H, pp(h, 0), s2)[0] - logpdf(xm, ym, H, pp(-h, 0), s2)[0]) / (2 * h)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_truth_mp.npz")
    np.savez(out, logpdf=float(lp), post_mean=np.array([float(v) for v in mean]), post_var=np.array([float(v) for v in var]),
             dlogpdf_dsigma2=float(d_s2), dlogpdf_dinv_lengthscale0=float(d_s), dlogpdf_dvariance0=float(d_v),
             logpdf_str=mp.nstr(lp, 30), digits=40)
    print("wrote", out, "logpdf =", mp.nstr(lp, 30), "d/dσ² =", mp.nstr(d_s2, 20), "d/ds0 =", mp.nstr(d_s, 20), "d/dv0 =", mp.nstr(d_v, 20))


if __name__ == "__main__":
    main()
