#!/usr/bin/env python
"""Generates tests/golden/c1_oilmm.npz: inputs and oracle outputs for BASELINE config 1 (OILMM
p=3, m=2, N=50, SEKernel + Matern32, sigma2=0.1) plus an IndependentMOGP and a general-ILMM case.

The reference (Julia) cannot run in the build image and its tests hold no numeric golden vectors
(SURVEY.md §8c), so these vectors come from the CPU oracle (`oracle/lmm_oracle.py`), cross-checked
at generation time against an independent extended-precision dense evaluation (np.longdouble
Cholesky of the pN x pN covariance written out below).  They pin the oracle and the CUDA path
against drift; they are NOT reference-generated.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import lmm_oracle as o  # noqa: E402


def longdouble_mvn_logpdf(C, y, mean):
    """Dense MVN logpdf in np.longdouble with a hand-written Cholesky (independent of LAPACK)."""
    C = np.array(C, dtype=np.longdouble)
    n = C.shape[0]
    L = np.zeros_like(C)
    for j in range(n):
        d = C[j, j] - np.dot(L[j, :j], L[j, :j])
        L[j, j] = np.sqrt(d)
        for i in range(j + 1, n):
            L[i, j] = (C[i, j] - np.dot(L[i, :j], L[j, :j])) / L[j, j]
    r = np.array(y, dtype=np.longdouble) - np.array(mean, dtype=np.longdouble)
    z = np.zeros(n, dtype=np.longdouble)
    for i in range(n):
        z[i] = (r[i] - np.dot(L[i, :i], z[:i])) / L[i, i]
    return float(-0.5 * (n * np.log(2 * np.longdouble(np.pi)) + 2 * np.sum(np.log(np.diag(L))) + np.dot(z, z)))


def main():
    rng = np.random.default_rng(20240416)
    N, p, m, Ns = 50, 3, 2, 7
    x = np.sort(rng.uniform(0, 5, N))
    xs = rng.uniform(0, 5, Ns)
    U, S = o.orthogonal_from_seed(p, m, seed=1)
    fs = [o.GP(o.Kernel(o.SE)), o.GP(o.Kernel(o.MATERN32))]
    y = rng.standard_normal(p * N)
    model = o.OILMMModel(fs, U, S)
    terms, reg = o.oilmm_logpdf_terms(model, x, 0.1, y)
    lp = float(np.sum(terms) + reg)
    # independent extended-precision check: dense pN x pN MVN (direct-difference distances)
    C = o.dense_mogp_cov(fs, model.H, x, form="direct") + 0.1 * np.eye(p * N)
    lp_ld = longdouble_mvn_logpdf(C, y, o.dense_mogp_mean(fs, model.H, x))
    assert abs(lp - lp_ld) < 1e-11 * abs(lp_ld), (lp, lp_ld)
    post = o.oilmm_posterior(model, x, 0.1, y)
    M, V = o.oilmm_mean_and_var(post, xs, 0.1)
    ys = rng.standard_normal(p * Ns)
    lp_post = o.oilmm_logpdf(post, xs, 0.1, ys)
    # IndependentMOGP with constant means (test/independent_mogp.jl:33-34)
    fs_i = [o.GP(o.Kernel(o.MATERN32), 30.0), o.GP(o.Kernel(o.SE, 0.5), 10.0)]
    y_i = np.concatenate([30 + rng.standard_normal(N), 10 + rng.standard_normal(N)])
    lp_i = o.imogp_logpdf(fs_i, x, 0.1, y_i)
    Mi, Vi = o.imogp_mean_and_var(o.imogp_posterior(fs_i, x, 0.1, y_i), xs, 0.1)
    # general ILMM
    Hg = rng.uniform(0, 1, (p, m))
    lp_g = o.ilmm_logpdf(fs, Hg, x, 0.1, y)
    Mg, Vg = o.ilmm_mean_and_var(o.ilmm_posterior(fs, Hg, x, 0.1, y), Hg, xs, 0.1)
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "c1_oilmm.npz")
    np.savez(out, x=x, xs=xs, U=U, S=S, y=y, ys=ys, sigma2=0.1, lml_terms=terms, regulariser=reg, logpdf=lp, logpdf_longdouble=lp_ld,
             post_mean=M, post_var=V, post_logpdf=lp_post, imogp_y=y_i, imogp_logpdf=lp_i, imogp_post_mean=Mi, imogp_post_var=Vi,
             ilmm_H=Hg, ilmm_logpdf=lp_g, ilmm_post_mean=Mg, ilmm_post_var=Vg)
    print("wrote", out, "logpdf", lp, "longdouble", lp_ld)


if __name__ == "__main__":
    main()
