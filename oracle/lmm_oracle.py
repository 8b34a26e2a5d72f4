"""CPU oracle for the LinearMixingModels.jl inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``linearmixingmodels.jl_b200/``,
``liblmm.so``) may import, call, link or execute this file; only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` do.

PARITY UNPINNED (in the strict sense): Julia is not available in the build container and the
reference's own tests hold no golden logpdf / mean / var numbers -- only
*equivalence identities* (OILMM == ILMM == dense multi-output GP == sum of single GPs) and a few
known answers (permutations, ``noise_var``, ``reshape_y``, ``Orthogonal`` validation).  This file
is a NumPy/SciPy Float64 restatement (LAPACK ``dpotrf``/``dtrtrs`` through OpenBLAS -- the routine
family Julia's ``cholesky`` and ``\\`` reach) pinned by exactly those identities and known answers
(``tests/test_oracle_identities.py``), by a long-double/dense cross-check and by 40-digit mpmath values of BASELINE
config 1 computed from the dense multi-output-GP definition with no shared code (``tests/golden/make_golden_mp.py``,
``tests/golden/c1_truth_mp.npz``: logpdf, posterior marginals, three logpdf derivatives) and of a general ILMM with
2-D inputs, ARD, Matern52 / Exponential / RationalQuadratic latents and constant means (``make_golden_mp2.py``,
``c2_truth_mp.npz``).

Every function cites the reference lines it follows (paths relative to /root/reference).  The
arithmetic of AbstractGPs 0.3.x / KernelFunctions 0.10.x / Distances 0.10.x is not vendored in the
reference; it is restated from the published behaviour of those packages at the versions pinned in
``examples/Manifest.toml`` and anchored on the reference's call sites.

Conventions: all arrays Float64.  ``x`` is (N,) or (N, D) (one row per input).  Multi-output
vectors are *by outputs*: ``y[(j-1)N + i]`` = output j at input i, i.e. ``y.reshape(p, N)`` in
NumPy C-order is the p x N matrix ``Y`` of ``reshape_y`` (src/ilmm.jl:43).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np
import scipy.linalg as sla

LOG2PI = math.log(2.0 * math.pi)

SE, MATERN32, MATERN52, EXPONENTIAL, RATQUAD, PERIODIC = 0, 1, 2, 3, 4, 5
COMPOSE_NONE, COMPOSE_SUM, COMPOSE_PRODUCT = 0, 1, 2
KERNEL_NAMES = {SE: "SEKernel", MATERN32: "Matern32Kernel", MATERN52: "Matern52Kernel", EXPONENTIAL: "ExponentialKernel",
                RATQUAD: "RationalQuadraticKernel", PERIODIC: "PeriodicKernel"}


# --------------------------------------------------------------------------------------------
# KernelFunctions / Distances semantics
# --------------------------------------------------------------------------------------------
@dataclass(frozen=True)
class Kernel:
    """``variance * (base ∘ ScaleTransform(inv_lengthscale))``.

    ``variance`` is KernelFunctions' ``ScaledKernel`` (``0.5 * SEKernel()``,
    test/independent_mogp.jl:108); ``inv_lengthscale`` its ``ScaleTransform`` -- inputs are
    multiplied by it *before* pairwise distances are taken.
    """

    kind: int = SE
    variance: float = 1.0
    inv_lengthscale: float = 1.0
    ard: Optional[tuple] = None  # KernelFunctions ``ARDTransform(v)``: x -> v .* x, composed with the ScaleTransform
    param: float = 1.0  # α of RationalQuadraticKernel, r of PeriodicKernel
    # KernelFunctions ``k1 + k2`` (KernelSum) / ``k1 * k2`` (KernelProduct): this kernel is term 0, ``terms`` the further
    # single kernels; kernelmatrix(::KernelSum) = sum of the components' kernel matrices, (::KernelProduct) their
    # elementwise product -- each component with its own transformed inputs and pairwise distances.
    op: int = COMPOSE_NONE
    terms: tuple = ()


@dataclass(frozen=True)
class GP:
    """``GP(mean_const, kernel)`` -- AbstractGPs ``GP(c::Real, k)`` / ``GP(k)`` (zero mean)."""

    kernel: Kernel = Kernel()
    mean_const: float = 0.0


def _as2d(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float64)
    return x.reshape(-1, 1) if x.ndim == 1 else x


def pairwise_sqdist(a: np.ndarray, b: Optional[np.ndarray] = None, *, form: str = "gemm") -> np.ndarray:
    """Squared Euclidean pairwise distances.

    ``form="gemm"`` restates Distances.jl 0.10 ``pairwise(SqEuclidean(), a, b; dims=1)``:
    ``max(|a_i|^2 + |b_j|^2 - 2 a_i.b_j, 0)`` with the cross term from a GEMM, and an exactly zero
    diagonal in the symmetric case.  ``form="direct"`` is ``sum((a_i - b_j)^2)`` (used to report
    the oracle-vs-oracle spread).
    """
    a = _as2d(a)
    sym = b is None
    bb = a if sym else _as2d(b)
    if form == "direct":
        d2 = np.zeros((a.shape[0], bb.shape[0]))
        for k in range(a.shape[1]):
            diff = a[:, k][:, None] - bb[:, k][None, :]
            d2 += diff * diff
        return d2
    sa = np.sum(a * a, axis=1)
    sb = sa if sym else np.sum(bb * bb, axis=1)
    if a.shape[1] == 1:
        # K=1 GEMM: a single correctly rounded product per entry (what dgemm/FMA-from-zero gives).
        r = a[:, 0][:, None] * bb[:, 0][None, :]
    else:
        r = a @ bb.T
    d2 = np.maximum((sa[:, None] + sb[None, :]) - 2.0 * r, 0.0)
    if sym:
        np.fill_diagonal(d2, 0.0)
    return d2


def kappa(kind: int, d2: np.ndarray, param: float = 1.0) -> np.ndarray:
    """Base kernel as a function of the squared distance (KernelFunctions ``kappa``).

    SE: exp(-d²/2) on SqEuclidean; Matern32: (1+√3 d)exp(-√3 d); Matern52:
    (1+√5 d+5d²/3)exp(-√5 d) on Euclidean d = sqrt(d²).
    """
    if kind == SE:
        return np.exp(-d2 / 2.0)
    d = np.sqrt(d2)
    if kind == MATERN32:
        s = math.sqrt(3.0) * d
        return (1.0 + s) * np.exp(-s)
    if kind == MATERN52:
        s = math.sqrt(5.0) * d
        return (1.0 + s + 5.0 * (d * d) / 3.0) * np.exp(-s)  # Julia: 5 * d^2 / 3
    if kind == EXPONENTIAL:  # KernelFunctions ExponentialKernel (= Matern12Kernel): exp(-d), metric Euclidean
        return np.exp(-d)
    if kind == RATQUAD:  # RationalQuadraticKernel(α): (1 + d²/(2α))^(-α), metric SqEuclidean
        return (1.0 + d2 / (2.0 * param)) ** (-param)
    raise ValueError(f"unsupported kernel kind {kind}")


def _sinpi(d: np.ndarray) -> np.ndarray:
    """sin(pi d) with the argument reduced first (Julia ``sinpi``): exact zeros at integers, no loss for large |d|."""
    r = d - 2.0 * np.round(d / 2.0)  # in [-1, 1]
    return np.sin(np.pi * r)


def _kernelmatrix_single(k: Kernel, x: np.ndarray, x2: Optional[np.ndarray], form: str) -> np.ndarray:
    sc = k.inv_lengthscale if k.ard is None else k.inv_lengthscale * np.asarray(k.ard, dtype=np.float64)[None, :]
    xs = _as2d(x) * sc
    x2s = None if x2 is None else _as2d(x2) * sc
    if k.kind == PERIODIC:
        # KernelFunctions PeriodicKernel(r): metric Distances.Sinus(r) = sum(abs2(sinpi(a_k - b_k) / r_k)), kappa(d) = exp(-d/2)
        bb = xs if x2s is None else x2s
        d = np.zeros((xs.shape[0], bb.shape[0]))
        for c in range(xs.shape[1]):
            sn = _sinpi(xs[:, c][:, None] - bb[:, c][None, :]) / k.param
            d += sn * sn
        return k.variance * np.exp(-0.5 * d)
    return k.variance * kappa(k.kind, pairwise_sqdist(xs, x2s, form=form), k.param)


def kernelmatrix(k: Kernel, x: np.ndarray, x2: Optional[np.ndarray] = None, *, form: str = "gemm") -> np.ndarray:
    """``kernelmatrix(k, x[, x2])`` = ``variance * map(κ, pairwise(metric, s*x, s*x2))``; for a KernelSum / KernelProduct the
    components' matrices added / multiplied left to right."""
    K = _kernelmatrix_single(k, x, x2, form)
    for t in k.terms:
        Kt = _kernelmatrix_single(t, x, x2, form)
        K = K * Kt if k.op == COMPOSE_PRODUCT else K + Kt
    return K


def kernel_kdiag(k: Kernel) -> float:
    v = k.variance
    for t in k.terms:
        v = v * t.variance if k.op == COMPOSE_PRODUCT else v + t.variance
    return v


def kernelmatrix_diag(k: Kernel, x: np.ndarray) -> np.ndarray:
    """``kernelmatrix_diag(k, x)`` = κ(0)·variance per point (summed / multiplied over the terms of a composite kernel)."""
    return np.full(_as2d(x).shape[0], kernel_kdiag(k), dtype=np.float64)


# --------------------------------------------------------------------------------------------
# AbstractGPs single-output exact GP
# --------------------------------------------------------------------------------------------
def _chol_lower(C: np.ndarray) -> np.ndarray:
    """``cholesky(Symmetric(C))`` -> lower factor (LAPACK dpotrf); raises LinAlgError if not PD."""
    return sla.cholesky(C, lower=True, check_finite=False)


def _fwd(L: np.ndarray, B: np.ndarray) -> np.ndarray:
    """``C.U' \\ B`` (dtrtrs, lower, no-trans)."""
    return sla.solve_triangular(L, B, lower=True, check_finite=False)


def _bwd(L: np.ndarray, B: np.ndarray) -> np.ndarray:
    return sla.solve_triangular(L, B, lower=True, trans="T", check_finite=False)


def gp_logpdf(f: GP, x, noise: float, y: np.ndarray, *, form: str = "gemm") -> float:
    """AbstractGPs ``logpdf(f(x, σ²), y)``: -(n log2π + logdet C + |C.U'⁻¹(y-m)|²)/2."""
    n = len(y)
    C = kernelmatrix(f.kernel, x, form=form)
    C[np.diag_indices_from(C)] += noise
    L = _chol_lower(C)
    z = _fwd(L, np.asarray(y, dtype=np.float64) - f.mean_const)
    return -0.5 * (n * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))


@dataclass
class PosteriorGP:
    """AbstractGPs ``PosteriorGP(prior, (α, C, x, δ))`` with C kept as its lower factor L."""

    prior: GP
    alpha: np.ndarray
    L: np.ndarray
    x: np.ndarray
    delta: np.ndarray
    form: str = "gemm"


def gp_posterior(f: GP, x, noise: float, y: np.ndarray, *, form: str = "gemm") -> PosteriorGP:
    """AbstractGPs ``posterior(f(x, σ²), y)``: C = chol(K+σ²I), δ = y-m, α = C \\ δ."""
    C = kernelmatrix(f.kernel, x, form=form)
    C[np.diag_indices_from(C)] += noise
    L = _chol_lower(C)
    delta = np.asarray(y, dtype=np.float64) - f.mean_const
    alpha = _bwd(L, _fwd(L, delta))
    return PosteriorGP(f, alpha, L, np.asarray(x, dtype=np.float64), delta, form)


def gp_mean(f, xs) -> np.ndarray:
    n = _as2d(xs).shape[0]
    if isinstance(f, PosteriorGP):
        Ksx = kernelmatrix(f.prior.kernel, xs, f.x, form=f.form)
        return f.prior.mean_const + Ksx @ f.alpha
    return np.full(n, f.mean_const, dtype=np.float64)


def gp_var(f, xs) -> np.ndarray:
    if isinstance(f, PosteriorGP):
        Kxs = kernelmatrix(f.prior.kernel, f.x, xs, form=f.form)
        V = _fwd(f.L, Kxs)
        return kernelmatrix_diag(f.prior.kernel, xs) - np.sum(V * V, axis=0)
    return kernelmatrix_diag(f.kernel, xs)


def gp_cov(f, xs) -> np.ndarray:
    if isinstance(f, PosteriorGP):
        Kxs = kernelmatrix(f.prior.kernel, f.x, xs, form=f.form)
        V = _fwd(f.L, Kxs)
        return kernelmatrix(f.prior.kernel, xs, form=f.form) - V.T @ V
    return kernelmatrix(f.kernel, xs)


def gp_cross_cov(f, xa, xb) -> np.ndarray:
    """AbstractGPs ``cov(f, x, x')``: prior kernel matrix, or K(x,x') - V_x' V_x' for a PosteriorGP."""
    if isinstance(f, PosteriorGP):
        Va = _fwd(f.L, kernelmatrix(f.prior.kernel, f.x, xa, form=f.form))
        Vb = _fwd(f.L, kernelmatrix(f.prior.kernel, f.x, xb, form=f.form))
        return kernelmatrix(f.prior.kernel, xa, xb, form=f.form) - Va.T @ Vb
    return kernelmatrix(f.kernel, xa, xb)


def gp_condition_again_marginals(post: "PosteriorGP", x2, noise2: float, y2: np.ndarray, xt) -> Tuple[np.ndarray, np.ndarray]:
    """Marginals at xt of ``posterior(post(x2, noise2), y2)`` by the textbook Gaussian update applied
    to the FIRST posterior (independent of the union-of-data route the library takes)."""
    C22 = gp_cov(post, x2)
    C22[np.diag_indices_from(C22)] += noise2
    L = _chol_lower(C22)
    Ct2 = gp_cross_cov(post, xt, x2)
    a = _bwd(L, _fwd(L, np.asarray(y2, dtype=np.float64) - gp_mean(post, x2)))
    V = _fwd(L, Ct2.T)
    return gp_mean(post, xt) + Ct2 @ a, gp_var(post, xt) - np.sum(V * V, axis=0)


def finite_marginals(f, xs, noise: float = 1e-18) -> Tuple[np.ndarray, np.ndarray]:
    """``mean_and_var(f(x, σ²))`` = (m(x), diag K + σ²); default FiniteGP noise is 1e-18."""
    return gp_mean(f, xs), gp_var(f, xs) + noise


def finite_logpdf(f, xs, noise: float, y: np.ndarray) -> float:
    """logpdf of a FiniteGP over a prior *or* posterior latent (used for ``logpdf(post(x*,σ²), y*)``)."""
    if not isinstance(f, PosteriorGP):
        return gp_logpdf(f, xs, noise, y)
    n = len(y)
    C = gp_cov(f, xs)
    C[np.diag_indices_from(C)] += noise
    L = _chol_lower(C)
    z = _fwd(L, np.asarray(y, dtype=np.float64) - gp_mean(f, xs))
    return -0.5 * (n * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))


def finite_rand(f, xs, noise: float, z: np.ndarray) -> np.ndarray:
    """``rand(rng, f(x, σ²))`` = m + C.U' * z with z ~ N(0, I) supplied by the caller."""
    C = gp_cov(f, xs)
    C[np.diag_indices_from(C)] += noise
    return gp_mean(f, xs) + _chol_lower(C) @ z


# --------------------------------------------------------------------------------------------
# LinearMixingModels: Orthogonal, project, regulariser (src/orthogonal_matrix.jl, src/oilmm.jl, src/ilmm.jl)
# --------------------------------------------------------------------------------------------
def validate_orthogonal(U: np.ndarray) -> None:
    """src/orthogonal_matrix.jl:21-23 -- ``isapprox(U'U, I)``.  Against a ``UniformScaling`` Julia's ``isapprox`` uses the
    operator 2-norm with ``|I| = 1``: ``opnorm(U'U - I) <= sqrt(eps) * max(opnorm(U'U), 1)``."""
    U = np.asarray(U, dtype=np.float64)
    m = U.shape[1]
    G = U.T @ U
    rtol = math.sqrt(np.finfo(np.float64).eps)
    if not (np.all(np.isfinite(G)) and np.linalg.norm(G - np.eye(m), 2) <= rtol * max(np.linalg.norm(G, 2), 1.0)):
        raise ValueError("`U` is not an orthogonal matrix")


def noise_var_known_answer() -> int:
    """test/ilmm.jl:56 -- ``noise_var(Diagonal(Fill(2, 3))) == 2``."""
    return 2


def reshape_y(y: np.ndarray, N: int) -> np.ndarray:
    """src/ilmm.jl:43 -- ``reshape(y, N, :)'`` -> p x N with Y[j, i] = y[j*N + i] (0-based)."""
    y = np.asarray(y, dtype=np.float64)
    return y.reshape(-1, N)


def project_orthogonal(U: np.ndarray, S: np.ndarray, sigma2: float) -> Tuple[np.ndarray, np.ndarray]:
    """src/oilmm.jl:20-30 -- T = sqrt(S) \\ U', ΣT = diag(σ² inv(S))."""
    S = np.asarray(S, dtype=np.float64)
    T = U.T / np.sqrt(S)[:, None]
    return T, sigma2 * (1.0 / S)


def regulariser_orthogonal(U: np.ndarray, S: np.ndarray, sigma2: float, Y: np.ndarray) -> float:
    """src/oilmm.jl:101-113."""
    n = Y.shape[1]
    p, m = U.shape
    R = (np.eye(p) - U @ U.T) @ Y
    return -(n * (float(np.sum(np.log(S))) + (p - m) * math.log(2.0 * math.pi * sigma2)) + float(np.sum(R * R)) / sigma2) / 2.0


def project_general(H: np.ndarray, sigma2: float) -> Tuple[np.ndarray, np.ndarray]:
    """src/ilmm.jl:61-68 -- includes the hard-coded 1e-9 jitter."""
    H = np.asarray(H, dtype=np.float64)
    m = H.shape[1]
    ST_inv = H.T @ H / sigma2 + 1e-9 * np.eye(m)
    L = _chol_lower(ST_inv)
    T = _bwd(L, _fwd(L, H.T)) / sigma2
    ST = T @ (sigma2 * np.eye(H.shape[0])) @ T.T
    return T, ST


def regulariser_general(H: np.ndarray, sigma2: float, Y: np.ndarray) -> float:
    """src/ilmm.jl:171-181."""
    p, m = H.shape
    n = Y.shape[1]
    T, ST = project_general(H, sigma2)
    _, logdet_ST = np.linalg.slogdet(ST)
    R = Y - H @ (T @ Y)
    return -(n * ((p - m) * LOG2PI + (p * math.log(sigma2) - logdet_ST)) + float(np.sum(R * R)) / sigma2) / 2.0


# --------------------------------------------------------------------------------------------
# IndependentMOGP (src/independent_mogp.jl)
# --------------------------------------------------------------------------------------------
def indices_outputs_to_features(N: int, p: int) -> np.ndarray:
    """src/independent_mogp.jl:135-139, 0-based.  Known answer test/independent_mogp.jl:86-98."""
    return np.arange(N * p).reshape(p, N).T.reshape(-1)


def indices_features_to_outputs(N: int, p: int) -> np.ndarray:
    """src/independent_mogp.jl:141-145, 0-based."""
    return np.arange(N * p).reshape(N, p).T.reshape(-1)


def imogp_logpdf(fs: Sequence[GP], x, sigma2: float, y: np.ndarray, *, form: str = "gemm") -> float:
    """src/independent_mogp.jl:74-80 -- y by outputs; sum of single-output logpdfs."""
    N = _as2d(x).shape[0]
    Y = reshape_y(y, N)
    return float(sum(gp_logpdf(f, x, sigma2, Y[i], form=form) for i, f in enumerate(fs)))


def imogp_logpdf_by_features(fs, x, sigma2, y_feat) -> float:
    """src/independent_mogp.jl:222-229."""
    N = _as2d(x).shape[0]
    return imogp_logpdf(fs, x, sigma2, np.asarray(y_feat)[indices_features_to_outputs(N, len(fs))])


def imogp_posterior(fs: Sequence[GP], x, sigma2: float, y: np.ndarray, *, form: str = "gemm") -> List[PosteriorGP]:
    """src/independent_mogp.jl:119-126."""
    N = _as2d(x).shape[0]
    Y = reshape_y(y, N)
    return [gp_posterior(f, x, sigma2, Y[i], form=form) for i, f in enumerate(fs)]


def imogp_mean_and_var(fs, xs, sigma2: float) -> Tuple[np.ndarray, np.ndarray]:
    """src/independent_mogp.jl:50-57 + FiniteGP noise: mean / var concatenated by outputs."""
    M = np.concatenate([gp_mean(f, xs) for f in fs])
    V = np.concatenate([gp_var(f, xs) for f in fs]) + sigma2
    return M, V


def imogp_cov(fs, xs) -> np.ndarray:
    """src/independent_mogp.jl:60-63 -- dense block diagonal."""
    return sla.block_diag(*[gp_cov(f, xs) for f in fs])


def imogp_mean_and_cov(fs, xs, sigma2: float) -> Tuple[np.ndarray, np.ndarray]:
    """AbstractGPs generic ``mean_and_cov(fx)`` = (mean(f,x), cov(f,x) + Σy) on an IndependentMOGP."""
    C = imogp_cov(fs, xs)
    return np.concatenate([gp_mean(f, xs) for f in fs]), C + sigma2 * np.eye(C.shape[0])


def imogp_rand(fs, xs, sigma2: float, z: np.ndarray) -> np.ndarray:
    """src/independent_mogp.jl:83-86 -- z holds N normals per latent, latent-major."""
    N = _as2d(xs).shape[0]
    Z = np.asarray(z, dtype=np.float64).reshape(len(fs), N)
    return np.concatenate([finite_rand(f, xs, sigma2, Z[i]) for i, f in enumerate(fs)])


# --------------------------------------------------------------------------------------------
# OILMM (src/oilmm.jl)
# --------------------------------------------------------------------------------------------
@dataclass
class OILMMModel:
    """``ILMM(IndependentMOGP(fs), Orthogonal(U, Diagonal(S)))`` -- fs may be priors or posteriors."""

    fs: list
    U: np.ndarray
    S: np.ndarray

    @property
    def H(self) -> np.ndarray:
        return self.U * np.sqrt(self.S)[None, :]


def _check_out_dim(p_x: int, p_H: int) -> None:
    if p_x != p_H:  # src/ilmm.jl:52
        raise RuntimeError("out dim of x != out dim of f.")


def oilmm_logpdf_terms(model: OILMMModel, x, sigma2: float, y: np.ndarray, *, form: str = "gemm"):
    """Per-latent lml terms and the regulariser -- src/oilmm.jl:79-93."""
    N = _as2d(x).shape[0]
    Y = reshape_y(y, N)
    _check_out_dim(Y.shape[0], model.U.shape[0])
    T, ST = project_orthogonal(model.U, model.S, sigma2)
    Ty = T @ Y
    lmls = [
        (gp_logpdf(f, x, ST[i], Ty[i], form=form) if not isinstance(f, PosteriorGP) else finite_logpdf(f, x, ST[i], Ty[i]))
        for i, f in enumerate(model.fs)
    ]
    return np.array(lmls), regulariser_orthogonal(model.U, model.S, sigma2, Y)


def oilmm_logpdf(model: OILMMModel, x, sigma2: float, y: np.ndarray, *, form: str = "gemm") -> float:
    lmls, reg = oilmm_logpdf_terms(model, x, sigma2, y, form=form)
    return float(np.sum(lmls) + reg)


def oilmm_posterior(model: OILMMModel, x, sigma2: float, y: np.ndarray, *, form: str = "gemm") -> OILMMModel:
    """src/oilmm.jl:116-134 -- the posterior is an OILMM whose latents are PosteriorGPs."""
    N = _as2d(x).shape[0]
    Y = reshape_y(y, N)
    _check_out_dim(Y.shape[0], model.U.shape[0])
    T, ST = project_orthogonal(model.U, model.S, sigma2)
    Ty = T @ Y
    posts = [gp_posterior(f, x, ST[i], Ty[i], form=form) for i, f in enumerate(model.fs)]
    return OILMMModel(posts, model.U, model.S)


def oilmm_mean_and_var(model: OILMMModel, xs, sigma2: float) -> Tuple[np.ndarray, np.ndarray]:
    """src/oilmm.jl:57-76 -- latent marginals (default FiniteGP noise 1e-18), mixed, by outputs."""
    ML = np.stack([finite_marginals(f, xs)[0] for f in model.fs])
    VL = np.stack([finite_marginals(f, xs)[1] for f in model.fs])
    H = model.H
    M = H @ ML
    V = (H * H) @ VL + sigma2
    return M.reshape(-1), V.reshape(-1)


def oilmm_rand(model: OILMMModel, xs, sigma2: float, z_latent: np.ndarray, z_noise: np.ndarray) -> np.ndarray:
    """src/oilmm.jl:40-54 -- latent draws use the default noise 1e-18; normals supplied by caller."""
    N = _as2d(xs).shape[0]
    Z = np.asarray(z_latent, dtype=np.float64).reshape(len(model.fs), N)
    X = np.stack([finite_rand(f, xs, 1e-18, Z[i]) for i, f in enumerate(model.fs)])  # m x N
    F = (model.H @ X).reshape(-1)
    return F + math.sqrt(sigma2) * np.asarray(z_noise, dtype=np.float64)


# --------------------------------------------------------------------------------------------
# ILMM with a general mixing matrix (src/ilmm.jl)
# --------------------------------------------------------------------------------------------
@dataclass
class ILMMPosterior:
    """``ILMM(PosteriorGP{IndependentMOGP}, H)`` -- joint (mN) posterior over the latents."""

    fs: list
    H: np.ndarray
    x: np.ndarray
    alpha: np.ndarray  # (mN,)
    L: np.ndarray  # (mN, mN) lower


def _latent_prior_cov(fs, x, form="gemm") -> np.ndarray:
    return sla.block_diag(*[kernelmatrix(f.kernel, x, form=form) for f in fs])


def ilmm_logpdf(fs: Sequence[GP], H: np.ndarray, x, sigma2: float, y: np.ndarray, *, form: str = "gemm") -> float:
    """src/ilmm.jl:150-163 -- projected (mN x mN) dense form + regulariser."""
    N = _as2d(x).shape[0]
    H = np.asarray(H, dtype=np.float64)
    p, m = H.shape
    Y = reshape_y(y, N)
    _check_out_dim(Y.shape[0], p)
    T, ST = project_general(H, sigma2)
    yproj = (T @ Y).reshape(-1)  # by outputs over the m latents
    C = _latent_prior_cov(fs, x, form) + np.kron(ST, np.eye(N))
    mean = np.concatenate([np.full(N, f.mean_const) for f in fs])
    L = _chol_lower(C)
    z = _fwd(L, yproj - mean)
    lml = -0.5 * (m * N * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))
    return lml + regulariser_general(H, sigma2, Y)


def ilmm_posterior(fs: Sequence[GP], H: np.ndarray, x, sigma2: float, y: np.ndarray, *, form: str = "gemm") -> ILMMPosterior:
    """src/ilmm.jl:184-198."""
    N = _as2d(x).shape[0]
    H = np.asarray(H, dtype=np.float64)
    Y = reshape_y(y, N)
    _check_out_dim(Y.shape[0], H.shape[0])
    T, ST = project_general(H, sigma2)
    yproj = (T @ Y).reshape(-1)
    C = _latent_prior_cov(fs, x, form) + np.kron(ST, np.eye(N))
    mean = np.concatenate([np.full(N, f.mean_const) for f in fs])
    L = _chol_lower(C)
    alpha = _bwd(L, _fwd(L, yproj - mean))
    return ILMMPosterior(list(fs), H, np.asarray(x, dtype=np.float64), alpha, L)


def _ilmm_latent_mean_and_cov(f, xs, form="gemm", jitter: float = 1e-18) -> Tuple[np.ndarray, np.ndarray]:
    """``mean_and_cov(f(x_mo, jitter))`` for prior (list of GP) or joint posterior; the default
    jitter is the FiniteGP default noise 1e-18 (src/ilmm.jl:115)."""
    Ns = _as2d(xs).shape[0]
    if isinstance(f, ILMMPosterior):
        m = len(f.fs)
        Ksx = sla.block_diag(*[kernelmatrix(g.kernel, xs, f.x, form=form) for g in f.fs])  # (mNs, mN)
        prior_mean = np.concatenate([np.full(Ns, g.mean_const) for g in f.fs])
        mean = prior_mean + Ksx @ f.alpha
        V = _fwd(f.L, Ksx.T)
        cov = sla.block_diag(*[kernelmatrix(g.kernel, xs, form=form) for g in f.fs]) - V.T @ V
    else:  # independent latents: priors (GP) or per-latent posteriors (PosteriorGP), src/independent_mogp.jl:50-63
        mean = np.concatenate([gp_mean(g, xs) for g in f])
        cov = sla.block_diag(*[gp_cov(g, xs) for g in f])
    return mean, cov + jitter * np.eye(cov.shape[0])


def ilmm_mean_and_cov(f, H: np.ndarray, xs, sigma2: float) -> Tuple[np.ndarray, np.ndarray]:
    """src/ilmm.jl:108-139 -- M = (H⊗I) m_lat, C = (H⊗I) C_lat (H⊗I)' + σ² I."""
    Ns = _as2d(xs).shape[0]
    mean, cov = _ilmm_latent_mean_and_cov(f, xs)
    Hf = np.kron(np.asarray(H, dtype=np.float64), np.eye(Ns))
    C = Hf @ cov @ Hf.T
    return Hf @ mean, C + sigma2 * np.eye(C.shape[0])


def ilmm_mean_and_var(f, H: np.ndarray, xs, sigma2: float) -> Tuple[np.ndarray, np.ndarray]:
    """src/ilmm.jl:122-129."""
    M, C = ilmm_mean_and_cov(f, H, xs, sigma2)
    return M, np.diag(C).copy()


def ilmm_rand(fs: Sequence[GP], H: np.ndarray, xs, sigma2: float, z_latent: np.ndarray, z_noise: np.ndarray) -> np.ndarray:
    """src/ilmm.jl:78-87 -- latent jitter 1e-12, then ``vec(reshape(latent, N, m) * H') + sqrt(σ²) ε``."""
    lat = imogp_rand(fs, xs, 1e-12, z_latent).reshape(len(fs), -1)  # m x N
    return (np.asarray(H, dtype=np.float64) @ lat).reshape(-1) + math.sqrt(sigma2) * np.asarray(z_noise, dtype=np.float64)


def ilmm_post_rand(post: ILMMPosterior, xs, sigma2: float, z_latent: np.ndarray, z_noise: np.ndarray) -> np.ndarray:
    """src/ilmm.jl:78-87 on ``ILMM(PosteriorGP{IndependentMOGP}, H)``: latent = m* + chol(C* + 1e-12 I) z."""
    mean, cov = _ilmm_latent_mean_and_cov(post, xs, jitter=1e-12)
    lat = (mean + _chol_lower(cov) @ np.asarray(z_latent, dtype=np.float64)).reshape(len(post.fs), -1)
    return (post.H @ lat).reshape(-1) + math.sqrt(sigma2) * np.asarray(z_noise, dtype=np.float64)


def ilmm_post_logpdf(post: ILMMPosterior, xs, sigma2: float, ys: np.ndarray) -> float:
    """src/ilmm.jl:150-163 on the posterior ILMM: logN(vec((TY*)') | m*, C* + ΣT ⊗ I) + regulariser."""
    Ns = _as2d(xs).shape[0]
    m = len(post.fs)
    Y = reshape_y(ys, Ns)
    T, ST = project_general(post.H, sigma2)
    yproj = (T @ Y).reshape(-1)
    mean, cov = _ilmm_latent_mean_and_cov(post, xs, jitter=0.0)
    L = _chol_lower(cov + np.kron(ST, np.eye(Ns)))
    z = _fwd(L, yproj - mean)
    lml = -0.5 * (m * Ns * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))
    return lml + regulariser_general(post.H, sigma2, Y)


def ilmm_condition_again_mean_and_var(post: ILMMPosterior, x2, sigma2_2: float, y2: np.ndarray, xt, sigma2_pred: float):
    """``mean_and_var(posterior(post(x2, σ2²), y2)(xt, σ²))`` for a general-ILMM posterior by the textbook
    update of the joint latent posterior (src/ilmm.jl:184-198 applied to PosteriorGP{IndependentMOGP} latents):
    the latent posterior GP is conditioned on vec((T2 Y2)') at x2 with noise ΣT2 ⊗ I, then mixed (src/ilmm.jl:122-129)."""
    N2, Nt = _as2d(x2).shape[0], _as2d(xt).shape[0]
    m = len(post.fs)
    H = post.H
    T2, ST2 = project_general(H, sigma2_2)
    Y2 = reshape_y(y2, N2)
    # joint latent posterior over [x2; xt] in latent-major order per block
    xa = np.concatenate([_as2d(x2), _as2d(xt)], axis=0)
    Na = N2 + Nt
    mean, cov = _ilmm_latent_mean_and_cov(post, xa if xa.shape[1] > 1 else xa[:, 0], jitter=0.0)
    i2 = np.concatenate([a * Na + np.arange(N2) for a in range(m)])
    it = np.concatenate([a * Na + N2 + np.arange(Nt) for a in range(m)])
    C22 = cov[np.ix_(i2, i2)] + np.kron(ST2, np.eye(N2))
    Ct2 = cov[np.ix_(it, i2)]
    L = _chol_lower(C22)
    d = (T2 @ Y2).reshape(-1) - mean[i2]
    mt = mean[it] + Ct2 @ _bwd(L, _fwd(L, d))
    V = _fwd(L, Ct2.T)
    Ctt = cov[np.ix_(it, it)] - V.T @ V + 1e-18 * np.eye(m * Nt)
    Hf = np.kron(H, np.eye(Nt))
    return Hf @ mt, np.diag(Hf @ Ctt @ Hf.T) + sigma2_pred


def _noise_matrix(Sigma, n: int) -> np.ndarray:
    S = np.asarray(Sigma, dtype=np.float64)
    return np.diag(S) if S.ndim == 1 else S


def imogp_logpdf_noise(fs: Sequence[GP], x, Sigma_y, y: np.ndarray) -> float:
    """AbstractGPs generic ``logpdf(f(x_mo, Σy), y)`` for an IndependentMOGP with Σy a vector (``Diagonal(v)``) or a dense
    matrix (test/independent_mogp.jl:72-75): MVN with mean(f, x), cov(f, x) + Σy (src/independent_mogp.jl:50-63)."""
    N = _as2d(x).shape[0]
    C = imogp_cov(fs, x) + _noise_matrix(Sigma_y, len(fs) * N)
    L = _chol_lower(C)
    z = _fwd(L, np.asarray(y, dtype=np.float64) - np.concatenate([gp_mean(f, x) for f in fs]))
    return -0.5 * (len(y) * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))


def imogp_posterior_noise_mean_and_cov(fs: Sequence[GP], x, Sigma_y, y: np.ndarray, xs, sigma2_pred: float):
    """AbstractGPs generic posterior of an IndependentMOGP under a vector / dense Σy, evaluated at (xs, σ²):
    textbook conditioning with the block-diagonal prior (src/independent_mogp.jl:60-71)."""
    N, Ns = _as2d(x).shape[0], _as2d(xs).shape[0]
    C = imogp_cov(fs, x) + _noise_matrix(Sigma_y, len(fs) * N)
    L = _chol_lower(C)
    delta = np.asarray(y, dtype=np.float64) - np.concatenate([gp_mean(f, x) for f in fs])
    alpha = _bwd(L, _fwd(L, delta))
    Ksx = sla.block_diag(*[kernelmatrix(f.kernel, xs, x) for f in fs])
    mean = np.concatenate([gp_mean(f, xs) for f in fs]) + Ksx @ alpha
    V = _fwd(L, Ksx.T)
    cov = imogp_cov(fs, xs) - V.T @ V + sigma2_pred * np.eye(len(fs) * Ns)
    return mean, cov


# --------------------------------------------------------------------------------------------
# Independent dense check: GP(LinearMixingModelKernel(kernels, H')) (test/ilmm.jl:5)
# --------------------------------------------------------------------------------------------
def dense_mogp_cov(fs: Sequence[GP], H: np.ndarray, x, x2=None, *, form: str = "gemm") -> np.ndarray:
    """Σ_i (h_i h_iᵀ) ⊗ K_i for by-outputs inputs (pN x pN')."""
    H = np.asarray(H, dtype=np.float64)
    out = None
    for i, f in enumerate(fs):
        blk = np.kron(np.outer(H[:, i], H[:, i]), kernelmatrix(f.kernel, x, x2, form=form))
        out = blk if out is None else out + blk
    return out


def dense_mogp_mean(fs: Sequence[GP], H: np.ndarray, x) -> np.ndarray:
    N = _as2d(x).shape[0]
    lat = np.stack([np.full(N, f.mean_const) for f in fs])
    return (np.asarray(H, dtype=np.float64) @ lat).reshape(-1)


def dense_mogp_logpdf(fs: Sequence[GP], H: np.ndarray, x, sigma2: float, y: np.ndarray, *, form: str = "gemm") -> float:
    C = dense_mogp_cov(fs, H, x, form=form)
    C[np.diag_indices_from(C)] += sigma2
    L = _chol_lower(C)
    z = _fwd(L, np.asarray(y, dtype=np.float64) - dense_mogp_mean(fs, H, x))
    return -0.5 * (len(y) * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))


def dense_mogp_posterior_mean_and_var(fs, H, x, sigma2, y, xs, sigma2_pred) -> Tuple[np.ndarray, np.ndarray]:
    """Textbook Gaussian conditioning on the dense pN x pN model; predictive noise added to var."""
    C = dense_mogp_cov(fs, H, x)
    C[np.diag_indices_from(C)] += sigma2
    L = _chol_lower(C)
    delta = np.asarray(y, dtype=np.float64) - dense_mogp_mean(fs, H, x)
    alpha = _bwd(L, _fwd(L, delta))
    Ksx = dense_mogp_cov(fs, H, xs, x)
    mean = dense_mogp_mean(fs, H, xs) + Ksx @ alpha
    V = _fwd(L, Ksx.T)
    var = np.diag(dense_mogp_cov(fs, H, xs)) - np.sum(V * V, axis=0) + sigma2_pred
    return mean, var


def missing_data_logpdf(fs: Sequence[GP], H: np.ndarray, x, sigma2: float, y: np.ndarray) -> float:
    """Heterotopic / missing-data ILMM (unsupported in the reference): entries of ``y`` that are NaN are
    unobserved; exact MVN logpdf of the observed entries under the dense model (test/ilmm.jl:5's ground-truth GP)."""
    y = np.asarray(y, dtype=np.float64)
    obs = np.flatnonzero(~np.isnan(y))
    C = dense_mogp_cov(fs, H, x)[np.ix_(obs, obs)] + sigma2 * np.eye(len(obs))
    L = _chol_lower(C)
    z = _fwd(L, y[obs] - dense_mogp_mean(fs, H, x)[obs])
    return -0.5 * (len(obs) * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(z @ z))


def missing_data_posterior_mean_and_var(fs, H, x, sigma2, y, xs, sigma2_pred) -> Tuple[np.ndarray, np.ndarray]:
    """Posterior marginals of ALL outputs at xs given the observed (non-NaN) entries of y: textbook conditioning of the dense model."""
    y = np.asarray(y, dtype=np.float64)
    obs = np.flatnonzero(~np.isnan(y))
    C = dense_mogp_cov(fs, H, x)[np.ix_(obs, obs)] + sigma2 * np.eye(len(obs))
    L = _chol_lower(C)
    alpha = _bwd(L, _fwd(L, y[obs] - dense_mogp_mean(fs, H, x)[obs]))
    Ksx = dense_mogp_cov(fs, H, xs, x)[:, obs]
    V = _fwd(L, Ksx.T)
    mean = dense_mogp_mean(fs, H, xs) + Ksx @ alpha
    var = np.diag(dense_mogp_cov(fs, H, xs)) - np.sum(V * V, axis=0) + sigma2_pred
    return mean, var


def missing_data_posterior_mean_and_cov(fs, H, x, sigma2, y, xs, sigma2_pred) -> Tuple[np.ndarray, np.ndarray]:
    """Posterior mean and full covariance (+ sigma2_pred I) of ALL outputs at xs given the observed entries of y (dense model)."""
    y = np.asarray(y, dtype=np.float64)
    obs = np.flatnonzero(~np.isnan(y))
    C = dense_mogp_cov(fs, H, x)[np.ix_(obs, obs)] + sigma2 * np.eye(len(obs))
    L = _chol_lower(C)
    alpha = _bwd(L, _fwd(L, y[obs] - dense_mogp_mean(fs, H, x)[obs]))
    Ksx = dense_mogp_cov(fs, H, xs, x)[:, obs]
    V = _fwd(L, Ksx.T)
    mean = dense_mogp_mean(fs, H, xs) + Ksx @ alpha
    cov = dense_mogp_cov(fs, H, xs) - V.T @ V
    cov[np.diag_indices_from(cov)] += sigma2_pred
    return mean, cov


# --------------------------------------------------------------------------------------------
# Gradients of the logpdf (for the rrule; checked against central finite differences in tests)
# --------------------------------------------------------------------------------------------
def _dkernel_ds(k: Kernel, x) -> np.ndarray:
    """d/d(inv_lengthscale) of kernelmatrix(k, x) (direct-difference distances; analytic)."""
    sc = k.inv_lengthscale if k.ard is None else k.inv_lengthscale * np.asarray(k.ard, dtype=np.float64)[None, :]
    xs = _as2d(x) * sc
    d2 = pairwise_sqdist(xs, form="direct")
    s = k.inv_lengthscale
    if k.kind == SE:
        return k.variance * (-np.exp(-d2 / 2.0) * d2 / s)
    d = np.sqrt(d2)
    if k.kind == MATERN32:
        return k.variance * (-3.0 * d2 * np.exp(-math.sqrt(3.0) * d) / s)
    if k.kind == MATERN52:
        return k.variance * (-(5.0 / 3.0) * d2 * (1.0 + math.sqrt(5.0) * d) * np.exp(-math.sqrt(5.0) * d) / s)
    if k.kind == EXPONENTIAL:
        return k.variance * (-d * np.exp(-d) / s)
    base = 1.0 + d2 / (2.0 * k.param)
    return k.variance * (-(d2 / s) * base ** (-k.param - 1.0))


def _dkernel_dard(k: Kernel, x) -> List[np.ndarray]:
    """d/d(ard_j) of kernelmatrix(k, x): dK/ds * (s / d²) * u_j² / ard_j with u = scaled coordinate differences (analytic)."""
    X = _as2d(x)
    if k.ard is None:
        return []
    a = np.asarray(k.ard, dtype=np.float64)
    xs = X * (k.inv_lengthscale * a[None, :])
    d2 = pairwise_sqdist(xs, form="direct")
    dKds = _dkernel_ds(k, x)
    with np.errstate(divide="ignore", invalid="ignore"):
        w = np.where(d2 > 0, dKds * k.inv_lengthscale / d2, 0.0)
    return [w * (xs[:, j][:, None] - xs[:, j][None, :]) ** 2 / a[j] for j in range(X.shape[1])]


def gp_logpdf_grad_ard(f: GP, x, noise: float, y: np.ndarray) -> np.ndarray:
    """d lml / d(ARD multipliers) of one latent (zeros if it has no ARDTransform)."""
    X = _as2d(x)
    if f.kernel.ard is None:
        return np.zeros(X.shape[1])
    C = kernelmatrix(f.kernel, x)
    C[np.diag_indices_from(C)] += noise
    L = _chol_lower(C)
    alpha = _bwd(L, _fwd(L, np.asarray(y, dtype=np.float64) - f.mean_const))
    G = 0.5 * (np.outer(alpha, alpha) - _bwd(L, _fwd(L, np.eye(C.shape[0]))))
    return np.array([float(np.sum(G * dK)) for dK in _dkernel_dard(f.kernel, x)])


def gp_logpdf_grad(f: GP, x, noise: float, y: np.ndarray):
    """(lml, d/dvariance, d/dinv_lengthscale, d/dmean, d/dnoise, d/dy) with G = (αα' - C⁻¹)/2."""
    C = kernelmatrix(f.kernel, x)
    n = C.shape[0]
    K = C.copy()
    C[np.diag_indices_from(C)] += noise
    L = _chol_lower(C)
    delta = np.asarray(y, dtype=np.float64) - f.mean_const
    alpha = _bwd(L, _fwd(L, delta))
    Cinv = _bwd(L, _fwd(L, np.eye(n)))
    G = 0.5 * (np.outer(alpha, alpha) - Cinv)
    lml = -0.5 * (n * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(delta @ alpha))
    return (lml, float(np.sum(G * K)) / f.kernel.variance, float(np.sum(G * _dkernel_ds(f.kernel, x))), float(np.sum(alpha)),
            float(np.trace(G)), -alpha)


def oilmm_logpdf_grad(model: OILMMModel, x, sigma2: float, y: np.ndarray):
    """Analytic gradient of src/oilmm.jl:79-93 w.r.t. latent hyper-parameters, σ² and y."""
    N = _as2d(x).shape[0]
    Y = reshape_y(y, N)
    p, m = model.U.shape
    T, ST = project_orthogonal(model.U, model.S, sigma2)
    Ty = T @ Y
    parts = [gp_logpdf_grad(f, x, ST[i], Ty[i]) for i, f in enumerate(model.fs)]
    R = (np.eye(p) - model.U @ model.U.T) @ Y
    resid = float(np.sum(R * R))
    lp = sum(q[0] for q in parts) + regulariser_orthogonal(model.U, model.S, sigma2, Y)
    g_sigma2 = sum(q[4] / model.S[i] for i, q in enumerate(parts)) - 0.5 * (N * (p - m) / sigma2 - resid / sigma2 ** 2)
    g_y = sum(np.outer(T[i], q[5]) for i, q in enumerate(parts)) - R / sigma2
    # mixing matrix: δ_i = U_i'Y/sqrt(S_i) - m_i, ν_i = σ²/S_i, regulariser -(n logdet S + |(I-UU')Y|²/σ²)/2
    S = model.S
    g_S = np.array([-(q[5] @ Ty[i]) / (2 * S[i]) - q[4] * sigma2 / S[i] ** 2 - N / (2 * S[i]) for i, q in enumerate(parts)])
    Z = model.U.T @ Y
    g_U = np.stack([Y @ q[5] / math.sqrt(S[i]) for i, q in enumerate(parts)], axis=1) + (R @ Z.T + Y @ (R.T @ model.U)) / sigma2
    return lp, {"variance": np.array([q[1] for q in parts]), "inv_lengthscale": np.array([q[2] for q in parts]),
                "mean_const": np.array([q[3] for q in parts]), "sigma2": float(g_sigma2), "y": g_y.reshape(-1), "U": g_U, "S": g_S}


def ilmm_logpdf_grad(fs: Sequence[GP], H: np.ndarray, x, sigma2: float, y: np.ndarray):
    """Analytic gradient of src/ilmm.jl:150-163 (projected form incl. `project` :61-68 and
    `regulariser` :171-181) w.r.t. latent hyper-parameters, σ², y and the dense mixing matrix H
    (what ``Zygote.gradient(logpdf, ilmmx, y)`` differentiates, test/ilmm.jl:31).

    With δ = vec((TY)') - μ, α = C⁻¹δ, G = (αα' - C⁻¹)/2 over the joint (mN) matrix
    C = blockdiag(K_a) + ΣT ⊗ I:  ∂/∂θ_a = <G_aa, ∂K_a/∂θ>,  ∂/∂ΣT[a,b] = Σ_n G[(a,n),(b,n)],
    ∂/∂(TY) = -A (A = α as m x N); the chain through T = M⁻¹H'/σ², M = H'H/σ² + 1e-9 I and
    ΣT = σ² T T' is done on the small matrices."""
    N = _as2d(x).shape[0]
    H = np.asarray(H, dtype=np.float64)
    p, m = H.shape
    Y = reshape_y(y, N)
    M = H.T @ H / sigma2 + 1e-9 * np.eye(m)
    W = np.linalg.inv(M)
    T = W @ H.T / sigma2
    ST = sigma2 * (T @ T.T)
    Z = T @ Y
    K = _latent_prior_cov(fs, x, "direct")
    C = K + np.kron(ST, np.eye(N))
    mean = np.concatenate([np.full(N, f.mean_const) for f in fs])
    L = _chol_lower(C)
    delta = Z.reshape(-1) - mean
    alpha = _bwd(L, _fwd(L, delta))
    Cinv = _bwd(L, _fwd(L, np.eye(m * N)))
    G = 0.5 * (np.outer(alpha, alpha) - Cinv)
    R = Y - H @ Z
    resid = float(np.sum(R * R))
    _, logdet_ST = np.linalg.slogdet(ST)
    lml = -0.5 * (m * N * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(delta @ alpha))
    reg = -(N * ((p - m) * LOG2PI + (p * math.log(sigma2) - logdet_ST)) + resid / sigma2) / 2.0
    A = alpha.reshape(m, N)
    g_var = np.zeros(m)
    g_s = np.zeros(m)
    g_ard = np.zeros((m, _as2d(x).shape[1]))
    for a, f in enumerate(fs):
        Gaa = G[a * N:(a + 1) * N, a * N:(a + 1) * N]
        g_var[a] = float(np.sum(Gaa * kernelmatrix(f.kernel, x, form="direct"))) / f.kernel.variance
        g_s[a] = float(np.sum(Gaa * _dkernel_ds(f.kernel, x)))
        for j, dK in enumerate(_dkernel_dard(f.kernel, x)):
            g_ard[a, j] = float(np.sum(Gaa * dK))
    B = np.array([[float(np.trace(G[a * N:(a + 1) * N, b * N:(b + 1) * N])) for b in range(m)] for a in range(m)])
    # direct cotangents
    HtR = H.T @ R
    Vm = HtR / sigma2 - A                       # m x N
    bar_T = Vm @ Y.T                            # -A Y' + H'R Y'/σ²
    bar_H = R @ Z.T / sigma2
    bar_ST = B + 0.5 * N * np.linalg.inv(ST)
    g_y = T.T @ Vm - R / sigma2
    g_sigma2 = -0.5 * (N * p / sigma2 - resid / sigma2 ** 2)
    # ΣT = σ² T T'
    g_sigma2 += float(np.sum(bar_ST * (T @ T.T)))
    bar_T = bar_T + sigma2 * (bar_ST + bar_ST.T) @ T
    # T = W H'/σ²
    g_sigma2 -= float(np.sum(bar_T * T)) / sigma2
    bar_H = bar_H + bar_T.T @ W / sigma2
    Zm = bar_T @ H / sigma2
    bar_M = -W @ Zm @ W
    bar_H = bar_H + H @ (bar_M + bar_M.T) / sigma2
    g_sigma2 -= float(np.sum(bar_M * (H.T @ H))) / sigma2 ** 2
    return lml + reg, {"variance": g_var, "inv_lengthscale": g_s, "mean_const": A.sum(axis=1), "sigma2": float(g_sigma2),
                       "y": g_y.reshape(-1), "H": bar_H, "ard": g_ard}


def oilmm_post_logpdf_grad(model: OILMMModel, xs, sigma2: float, ys: np.ndarray):
    """Gradient of ``logpdf(post(x*, σ²), y*)`` (test/oilmm.jl:32 `gradient(logpdf, po, y_test)`) w.r.t. σ² and
    y*, the posterior's data (α, C, x) and hyper-parameters held fixed.  ``model.fs`` are PosteriorGPs
    (OILMM) -- pass ``U = I, S = 1`` for an IndependentMOGP posterior (no regulariser: p == m)."""
    Ns = _as2d(xs).shape[0]
    Y = reshape_y(ys, Ns)
    p, m = model.U.shape
    T, ST = project_orthogonal(model.U, model.S, sigma2)
    Ty = T @ Y
    lp, g_s2, g_y = 0.0, 0.0, np.zeros((p, Ns))
    for i, f in enumerate(model.fs):
        C = gp_cov(f, xs) + ST[i] * np.eye(Ns)
        L = _chol_lower(C)
        delta = Ty[i] - gp_mean(f, xs)
        alpha = _bwd(L, _fwd(L, delta))
        Li = _fwd(L, np.eye(Ns))
        lp += -0.5 * (Ns * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(delta @ alpha))
        g_s2 += 0.5 * (float(alpha @ alpha) - float(np.sum(Li * Li))) / model.S[i]
        g_y -= np.outer(T[i], alpha)
    if p > m:
        R = (np.eye(p) - model.U @ model.U.T) @ Y
        resid = float(np.sum(R * R))
        lp += regulariser_orthogonal(model.U, model.S, sigma2, Y)
        g_s2 += -0.5 * (Ns * (p - m) / sigma2 - resid / sigma2 ** 2)
        g_y -= R / sigma2
    else:
        lp += -0.5 * Ns * float(np.sum(np.log(model.S)))
    return lp, {"sigma2": float(g_s2), "y": g_y.reshape(-1)}


def ilmm_post_logpdf_grad(post: ILMMPosterior, xs, sigma2: float, ys: np.ndarray):
    """Same for a general-ILMM posterior (test/ilmm.jl:32): σ² enters through `project` (T, ΣT) and the
    regulariser; the joint latent posterior (m*, C*) is held fixed."""
    Ns = _as2d(xs).shape[0]
    H = post.H
    p, m = H.shape
    Y = reshape_y(ys, Ns)
    M = H.T @ H / sigma2 + 1e-9 * np.eye(m)
    W = np.linalg.inv(M)
    T = W @ H.T / sigma2
    ST = sigma2 * (T @ T.T)
    Z = T @ Y
    mean, cov = _ilmm_latent_mean_and_cov(post, xs, jitter=0.0)
    C = cov + np.kron(ST, np.eye(Ns))
    L = _chol_lower(C)
    delta = Z.reshape(-1) - mean
    alpha = _bwd(L, _fwd(L, delta))
    Cinv = _bwd(L, _fwd(L, np.eye(m * Ns)))
    G = 0.5 * (np.outer(alpha, alpha) - Cinv)
    R = Y - H @ Z
    resid = float(np.sum(R * R))
    _, logdet_ST = np.linalg.slogdet(ST)
    lp = -0.5 * (m * Ns * LOG2PI + 2.0 * float(np.sum(np.log(np.diag(L)))) + float(delta @ alpha))
    lp += -(Ns * ((p - m) * LOG2PI + (p * math.log(sigma2) - logdet_ST)) + resid / sigma2) / 2.0
    A = alpha.reshape(m, Ns)
    B = np.array([[float(np.trace(G[a * Ns:(a + 1) * Ns, b * Ns:(b + 1) * Ns])) for b in range(m)] for a in range(m)])
    Vm = H.T @ R / sigma2 - A
    bar_T = Vm @ Y.T
    bar_ST = B + 0.5 * Ns * np.linalg.inv(ST)
    g_y = T.T @ Vm - R / sigma2
    g_s2 = -0.5 * (Ns * p / sigma2 - resid / sigma2 ** 2) + float(np.sum(bar_ST * (T @ T.T)))
    bar_T = bar_T + sigma2 * (bar_ST + bar_ST.T) @ T
    g_s2 -= float(np.sum(bar_T * T)) / sigma2
    bar_M = -W @ (bar_T @ H / sigma2) @ W
    g_s2 -= float(np.sum(bar_M * (H.T @ H))) / sigma2 ** 2
    return lp, {"sigma2": float(g_s2), "y": g_y.reshape(-1)}


# --------------------------------------------------------------------------------------------
# Synthetic workloads shared by tests and bench: identical bytes for oracle and GPU
# --------------------------------------------------------------------------------------------
def orthogonal_from_seed(p: int, m: int, seed: int = 1) -> Tuple[np.ndarray, np.ndarray]:
    """U, S = thin SVD of uniform(0,1) p x m, as the tests/notebook do (test/oilmm.jl:45-46)."""
    rng = np.random.default_rng(seed)
    U, S, _ = np.linalg.svd(rng.uniform(0.0, 1.0, size=(p, m)), full_matrices=False)
    return np.ascontiguousarray(U), np.ascontiguousarray(S)
