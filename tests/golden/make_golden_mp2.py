#!/usr/bin/env python
"""Generates tests/golden/c2_truth_mp.npz: 40-digit (mpmath) evaluation of a small GENERAL ILMM (dense, non-orthogonal H;
src/ilmm.jl) that exercises what config 1 does not: 2-D inputs, an ARDTransform, Matern52 / Exponential / RationalQuadratic
latents, constant means.  Written straight from the definition the reference's tests use as ground truth (test/ilmm.jl:5,
`LinearMixingModelKernel`):  y ~ N((H ⊗ I) m(x), Σ_i (h_i h_i') ⊗ K_i + σ² I), by outputs -- no code shared with
oracle/lmm_oracle.py.  Stored: inputs, logpdf, posterior marginals at x*, d logpdf / dσ², d logpdf / d(ARD multiplier 2 of
latent 1), d logpdf / dH[2,3] by 40-digit central differences.

    python tests/golden/make_golden_mp2.py        (about a minute)
"""
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 40


def kappa(kind, d2, alpha):
    if kind == "m52":
        d = mp.sqrt(d2)
        s = mp.sqrt(5) * d
        return (1 + s + 5 * d2 / 3) * mp.exp(-s)
    if kind == "exp":
        return mp.exp(-mp.sqrt(d2))
    return (1 + d2 / (2 * alpha)) ** (-alpha)  # rq


def latent_cov(lat, xa, xb):
    kind, var, s, ard, alpha, _ = lat
    out = mp.zeros(len(xa), len(xb))
    for a, pa in enumerate(xa):
        for b, pb in enumerate(xb):
            d2 = mp.fsum(((s * ard[k]) * (pa[k] - pb[k])) ** 2 for k in range(len(pa)))
            out[a, b] = var * kappa(kind, d2, alpha)
    return out


def cov_dense(xa, xb, H, lats, noise=None):
    p, m = len(H), len(H[0])
    na, nb = len(xa), len(xb)
    Ks = [latent_cov(l, xa, xb) for l in lats]
    C = mp.zeros(p * na, p * nb)
    for j in range(p):
        for j2 in range(p):
            for i in range(m):
                w = H[j][i] * H[j2][i]
                for a in range(na):
                    for b in range(nb):
                        C[j * na + a, j2 * nb + b] += w * Ks[i][a, b]
    if noise is not None:
        for k in range(p * na):
            C[k, k] += noise
    return C


def mean_dense(n, H, lats):
    return [mp.fsum(H[j][i] * lats[i][5] for i in range(len(lats))) for j in range(len(H)) for _ in range(n)]


def chol(C):
    n = C.rows
    L = mp.zeros(n, n)
    for j in range(n):
        L[j, j] = mp.sqrt(C[j, j] - mp.fsum(L[j, k] ** 2 for k in range(j)))
        for i in range(j + 1, n):
            L[i, j] = (C[i, j] - mp.fsum(L[i, k] * L[j, k] for k in range(j))) / L[j, j]
    return L


def fwd(L, b):
    z = [mp.mpf(0)] * L.rows
    for i in range(L.rows):
        z[i] = (b[i] - mp.fsum(L[i, k] * z[k] for k in range(i))) / L[i, i]
    return z


def bwd(L, b):
    n = L.rows
    z = [mp.mpf(0)] * n
    for i in reversed(range(n)):
        z[i] = (b[i] - mp.fsum(L[k, i] * z[k] for k in range(i + 1, n))) / L[i, i]
    return z


def logpdf(x, y, H, lats, s2):
    L = chol(cov_dense(x, x, H, lats, s2))
    mu = mean_dense(len(x), H, lats)
    r = [a - b for a, b in zip(y, mu)]
    z = fwd(L, r)
    n = len(y)
    return -(n * mp.log(2 * mp.pi) + 2 * mp.fsum(mp.log(L[i, i]) for i in range(n)) + mp.fsum(v * v for v in z)) / 2, L, r


def main():
    rng = np.random.default_rng(20241018)
    N, Ns, p, m, D = 16, 5, 4, 3, 2
    x = rng.uniform(0, 3, (N, D))
    xs = rng.uniform(0, 3, (Ns, D))
    Hf = rng.uniform(0.1, 1.0, (p, m))
    y = rng.standard_normal(p * N) + 0.3
    ard0 = np.array([0.7, 1.3])
    mpf = lambda a: mp.mpf(float(a))
    one2 = [mp.mpf(1), mp.mpf(1)]
    # (kind, variance, inverse lengthscale, ARD multipliers, alpha, constant mean) -- Float64 literals, exactly
    lats = [("m52", mpf(1.2), mpf(0.9), [mpf(v) for v in ard0], mp.mpf(1), mpf(0.5)),
            ("exp", mpf(0.8), mpf(1.4), one2, mp.mpf(1), mpf(-1.0)),
            ("rq", mpf(1.0), mpf(0.6), one2, mpf(1.7), mpf(0.0))]
    H = [[mpf(Hf[j, i]) for i in range(m)] for j in range(p)]
    xm = [[mpf(v) for v in row] for row in x]
    xsm = [[mpf(v) for v in row] for row in xs]
    ym = [mpf(v) for v in y]
    s2 = mpf(0.05)
    lp, L, r = logpdf(xm, ym, H, lats, s2)
    alpha = bwd(L, fwd(L, r))
    Ksx = cov_dense(xsm, xm, H, lats)
    Kss = cov_dense(xsm, xsm, H, lats)
    mus = mean_dense(Ns, H, lats)
    mean, var = [], []
    for q in range(p * Ns):
        row = [Ksx[q, c] for c in range(p * N)]
        mean.append(mus[q] + mp.fsum(a * b for a, b in zip(row, alpha)))
        v = fwd(L, row)
        var.append(Kss[q, q] - mp.fsum(t * t for t in v) + s2)
    h = mp.mpf(10) ** -12
    d_s2 = (logpdf(xm, ym, H, lats, s2 + h)[0] - logpdf(xm, ym, H, lats, s2 - h)[0]) / (2 * h)

    def with_ard(dv):
        l0 = lats[0]
        return [(l0[0], l0[1], l0[2], [l0[3][0], l0[3][1] + dv], l0[4], l0[5])] + lats[1:]

    d_ard = (logpdf(xm, ym, H, with_ard(h), s2)[0] - logpdf(xm, ym, H, with_ard(-h), s2)[0]) / (2 * h)

    def with_H(dv):
        H2 = [row[:] for row in H]
        H2[1][2] += dv
        return H2

    d_H = (logpdf(xm, ym, with_H(h), lats, s2)[0] - logpdf(xm, ym, with_H(-h), lats, s2)[0]) / (2 * h)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c2_truth_mp.npz")
    np.savez(out, x=x, xs=xs, H=Hf, y=y, ard0=ard0, sigma2=0.05,
             logpdf=float(lp), post_mean=np.array([float(v) for v in mean]), post_var=np.array([float(v) for v in var]),
             dlogpdf_dsigma2=float(d_s2), dlogpdf_dard0_1=float(d_ard), dlogpdf_dH12=float(d_H), logpdf_str=mp.nstr(lp, 30), digits=40)
    print("wrote", out, "logpdf =", mp.nstr(lp, 30), "d/dσ² =", mp.nstr(d_s2, 20), "d/dard =", mp.nstr(d_ard, 20), "d/dH =", mp.nstr(d_H, 20))


if __name__ == "__main__":
    main()
