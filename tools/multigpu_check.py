#!/usr/bin/env python
"""Multi-rank parity check (run under torchrun, one rank per GPU): OILMM logpdf / posterior /
mean_and_var / rand / posterior-logpdf / sweep with latents sharded over the ranks and the
in-library NCCL all-reduce, against the CPU oracle on identical inputs."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402
from oracle import lmm_oracle as o  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = lmm.Context(local)
    lmm.set_default_context(ctx)
    lmm.dist.init_context_distributed(ctx)
    rng = np.random.default_rng(0)
    N, p, m, Ns = 700, 8, 5, 33
    x = np.sort(rng.uniform(0, 7, N))
    xs = rng.uniform(0, 7, Ns)
    U, S = o.orthogonal_from_seed(p, m, seed=1)
    kinds = [o.SE, o.MATERN32, o.MATERN52, o.SE, o.MATERN32]
    fs = [o.GP(o.Kernel(kinds[i], float(rng.uniform(0.5, 1.5)), float(rng.uniform(0.5, 2.0))), float(rng.normal())) for i in range(m)]
    y = rng.standard_normal(p * N)
    om = o.OILMMModel(fs, U, S)
    names = {o.SE: lmm.SEKernel, o.MATERN32: lmm.Matern32Kernel, o.MATERN52: lmm.Matern52Kernel}
    gps = [lmm.GP(g.mean_const, (g.kernel.variance * names[g.kernel.kind]()).compose(lmm.ScaleTransform(g.kernel.inv_lengthscale))) for g in fs]
    f = lmm.ILMM(lmm.independent_mogp(gps), lmm.Orthogonal(U, S))
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    ref = o.oilmm_logpdf(om, x, 0.1, y)
    got = lmm.logpdf(fx, y)
    errs = {"logpdf": abs(got - ref) / abs(ref)}
    post, lp = lmm.posterior(fx, y, with_logpdf=True)
    errs["posterior_logpdf_same"] = abs(lp - got)
    M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
    opost = o.oilmm_posterior(om, x, 0.1, y)
    Mr, Vr = o.oilmm_mean_and_var(opost, xs, 0.1)
    errs["mean"] = float(np.max(np.abs(M - Mr) / (np.abs(Mr) + 1e-10)))
    errs["var"] = float(np.max(np.abs(V - Vr) / np.abs(Vr)))
    ys = np.random.default_rng(5).standard_normal(p * Ns)
    errs["post_logpdf"] = abs(lmm.logpdf(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.2), ys) - o.oilmm_logpdf(opost, xs, 0.2, ys)) / abs(
        o.oilmm_logpdf(opost, xs, 0.2, ys))
    scales = np.array([0.7, 1.0, 1.6])
    sw = lmm.logpdf_sweep(fx, y, scales)
    for s, v in zip(scales, sw):
        fs_s = [o.GP(o.Kernel(g.kernel.kind, g.kernel.variance, g.kernel.inv_lengthscale * s), g.mean_const) for g in fs]
        r = o.oilmm_logpdf(o.OILMMModel(fs_s, U, S), x, 0.1, y)
        errs[f"sweep_{s}"] = abs(v - r) / abs(r)
    # rand (Matern latents at few points so that K + 1e-18 I is numerically PD)
    xr = np.linspace(0, 10, 6)
    fr = [o.GP(o.Kernel(o.MATERN32, 1.0, 1.0 + 0.3 * i), 0.1 * i) for i in range(m)]
    gr = [lmm.GP(g.mean_const, lmm.Matern32Kernel().compose(lmm.ScaleTransform(g.kernel.inv_lengthscale))) for g in fr]
    frx = lmm.ILMM(lmm.independent_mogp(gr), lmm.Orthogonal(U, S))(lmm.MOInputIsotopicByOutputs(xr, p), 0.1)
    gq = np.random.default_rng(21)
    zl, zn = gq.standard_normal(m * 6), gq.standard_normal(p * 6)
    s = lmm.rand(np.random.default_rng(21), frx)
    errs["rand"] = float(np.max(np.abs(s - o.oilmm_rand(o.OILMMModel(fr, U, S), xr, 0.1, zl, zn))))
    # dense covariance, sequential conditioning and the logpdf gradient across ranks
    xc = np.sort(rng.uniform(0, 7, 9))
    Mc, Cc = lmm.mean_and_cov(post(lmm.MOInputIsotopicByOutputs(xc, p), 0.1))
    Mcr, Ccr = o.ilmm_mean_and_cov(opost.fs, om.H, xc, 0.1)
    errs["cov"] = float(np.max(np.abs(Cc - Ccr)) / np.max(np.abs(Ccr)))
    errs["cov_mean"] = float(np.max(np.abs(Mc - Mcr)))
    x2 = rng.uniform(0, 7, 21)
    y2 = rng.standard_normal(p * 21)
    post2 = lmm.posterior(post(lmm.MOInputIsotopicByOutputs(x2, p), 0.1), y2)
    xu = np.concatenate([x, x2])
    yu = np.concatenate([np.concatenate([y.reshape(p, N)[j], y2.reshape(p, 21)[j]]) for j in range(p)])
    Mu, Vu = o.oilmm_mean_and_var(o.oilmm_posterior(om, xu, 0.1, yu), xs, 0.1)
    M2, V2 = lmm.mean_and_var(post2(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
    errs["cond_mean"] = float(np.max(np.abs(M2 - Mu) / (np.abs(Mu) + 1e-8)))
    errs["cond_var"] = float(np.max(np.abs(V2 - Vu) / np.abs(Vu)))
    lpg, g = lmm.logpdf_and_gradient(fx, y, with_grad_y=True)
    lpr, gr = o.oilmm_logpdf_grad(om, x, 0.1, y)
    errs["grad_value"] = abs(lpg - lpr) / abs(lpr)
    for k in ("variance", "inv_lengthscale", "mean_const", "y", "S", "U"):
        errs["grad_" + k] = float(np.max(np.abs(g[k] - gr[k])) / (np.max(np.abs(gr[k])) + 1e-12))
    errs["grad_sigma2"] = abs(g["sigma2"] - gr["sigma2"]) / abs(gr["sigma2"])
    # posterior-predictive logpdf gradient and a per-observation-noise IndependentMOGP across ranks
    lpp, gp_ = lmm.logpdf_and_gradient(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.2), ys, with_grad_y=True)
    lppr, gpr = o.oilmm_post_logpdf_grad(opost, xs, 0.2, ys)
    errs["postgrad_value"] = abs(lpp - lppr) / abs(lppr)
    errs["postgrad_sigma2"] = abs(gp_["sigma2"] - gpr["sigma2"]) / abs(gpr["sigma2"])
    errs["postgrad_y"] = float(np.max(np.abs(gp_["y"] - gpr["y"])) / np.max(np.abs(gpr["y"])))
    fi = lmm.independent_mogp(gps)
    vn = np.random.default_rng(8).uniform(0.05, 0.5, m * N)
    yi = y[: m * N]
    errs["imogp_vector_noise"] = abs(lmm.logpdf(fi(lmm.MOInputIsotopicByOutputs(x, m), vn), yi) - o.imogp_logpdf_noise(fs, x, vn, yi)) / abs(
        o.imogp_logpdf_noise(fs, x, vn, yi))
    post2.f.fs[0]._owner.free()
    # a factorisation that fails on SOME ranks only: prior rand of an OILMM whose first latents are SE kernels on 700
    # points 0.01 apart (K + 1e-18 I is numerically singular: PosDefException, as in the reference) and whose last latents are
    # Exponential kernels (fine).  Every rank must raise the SAME exception (collective verdict) instead of some ranks hanging in
    # the all-reduce that follows, and the next collective call must still work.
    gmix = [lmm.GP(lmm.SEKernel()) if i < (m + 1) // 2 else lmm.GP(lmm.ExponentialKernel()) for i in range(m)]
    fmix = lmm.ILMM(lmm.independent_mogp(gmix), lmm.Orthogonal(U, S))
    verdict = [-1.0, -1.0]
    try:
        lmm.rand(np.random.default_rng(3), fmix(lmm.MOInputIsotopicByOutputs(x, p), 0.1))
    except lmm.PosDefException as e:
        verdict = [float(e.info), float(e.latent)]
    tv = torch.tensor(verdict, dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(tv) for _ in range(world)]
    dist.all_gather(allv, tv)
    same = all(bool(torch.equal(a, allv[0])) for a in allv)
    errs["posdef_collective"] = 0.0 if (same and verdict[0] > 0 and verdict[1] == 0.0) else 1.0
    errs["after_posdef_logpdf"] = abs(lmm.logpdf(fx, y) - ref) / abs(ref)
    worst = max(errs.values())
    ok = worst < 1e-6 and max(errs[k] for k in ("logpdf", "mean", "var", "post_logpdf", "grad_value")) < 1e-9
    print(f"rank {rank}/{world}: {'OK' if ok else 'FAIL'} " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()), flush=True)
    post.f.fs[0]._owner.free()
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
