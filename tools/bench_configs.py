#!/usr/bin/env python
"""Timings of the other BASELINE.json configs on one GPU (C2, C3, a 1-GPU slice of C5) -- these are
parity-test cases, not the headline bench line; printed as JSON lines for DESIGN.md / profiles."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402


def timed(fn, reps=3):
    fn()
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        out = fn()
        best = min(best, time.perf_counter() - t0)
    return best, out


def main():
    ctx = lmm.default_context()
    rng = np.random.default_rng(0)
    # ---- C2: ILMM p=8 m=4 N=2048
    N, p, m = 2048, 8, 4
    x = np.sort(rng.uniform(0, N / 100.0, N))
    H = np.random.default_rng(1).uniform(0, 1, (p, m))
    ks = [lmm.SEKernel(), lmm.Matern32Kernel(), lmm.Matern52Kernel(), lmm.SEKernel()]
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP(k) for k in ks]), H)
    y = rng.standard_normal(p * N)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    for form, dim in ((0, m * N), (1, p * N)):
        lmm.set_ilmm_form(form)
        t, _ = timed(lambda: lmm.logpdf(fx, y))
        tm = ctx.last_timings()
        print(json.dumps({"config": "C2 ILMM p=8 m=4 N=2048", "form": "projected mN" if form == 0 else "dense pN", "dim": dim,
                          "wall_ms": t * 1e3, "assemble_ms": tm[1], "chol_ms": tm[2], "chol_tflops": dim ** 3 / 3 / (tm[2] * 1e-3) / 1e12}), flush=True)
    lmm.set_ilmm_form(0)
    # ---- C3: OILMM p=64 m=16 N=8192 Matern52, logpdf + posterior + marginals at N*=1024
    N, p, m, Ns = 8192, 64, 16, 1024
    x = np.sort(rng.uniform(0, N / 100.0, N))
    xs = rng.uniform(0, N / 100.0, Ns)
    U, S, _ = np.linalg.svd(np.random.default_rng(1).uniform(0, 1, (p, m)), full_matrices=False)
    inv_ls = np.random.default_rng(2).uniform(0.5, 2.0, m)
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.Matern52Kernel().compose(lmm.ScaleTransform(float(s)))) for s in inv_ls]), lmm.Orthogonal(U, S))
    y = rng.standard_normal(p * N)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)

    def evalc3():
        post, lp = lmm.posterior(fx, y, with_logpdf=True)
        tm = ctx.last_timings().copy()
        t0 = time.perf_counter()
        M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), 0.1))
        tp = time.perf_counter() - t0
        tdev = float(ctx.last_timings()[5])  # CUDA-event time of the prediction call
        post.f.fs[0]._owner.free()
        return tm, tp, tdev

    t, (tm, tp, tdev) = timed(evalc3)
    print(json.dumps({"config": "C3 OILMM p=64 m=16 N=8192 Matern52", "eval_plus_marginals_wall_ms": t * 1e3, "logpdf_posterior_ms": tm[0],
                      "kmat_ms": tm[1], "chol_ms": tm[2], "solves_ms": tm[3], "chol_tflops": m * N ** 3 / 3 / (tm[2] * 1e-3) / 1e12,
                      "marginals_Ns1024_wall_ms": tp * 1e3, "marginals_Ns1024_device_ms": tdev,
                      "marginals_tflops": m * (N * N * Ns) / (tdev * 1e-3) / 1e12}), flush=True)
    # ---- gradient (rrule) at the C3 shape: value + d/d(hyper-parameters, σ², y)
    t, (lp, g) = timed(lambda: lmm.logpdf_and_gradient(fx, y, with_grad_y=True), reps=2)
    print(json.dumps({"config": "C3 shape: logpdf + gradient (batched potri + fused kernel-gradient reduction)", "wall_ms": t * 1e3,
                      "tflops": m * N ** 3 / t / 1e12, "note": "N^3 flop per latent = potrf + triangular TRSM + SYRK"}), flush=True)
    # ---- C5 slice on one GPU: p=256, m=16 of 128 latents, N=8192, 4 of 32 lengthscales
    N, p, m, nsw = 8192, 256, 16, 4
    x = np.sort(rng.uniform(0, N / 100.0, N))
    U, S, _ = np.linalg.svd(np.random.default_rng(1).uniform(0, 1, (p, m)), full_matrices=False)
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.SEKernel()) for _ in range(m)]), lmm.Orthogonal(U, S))
    y = rng.standard_normal(p * N)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    scales = np.geomspace(0.25, 4, 32)[::8]
    t, out = timed(lambda: lmm.logpdf_sweep(fx, y, scales), reps=2)
    print(json.dumps({"config": "C5 slice: OILMM p=256 m=16(of 128) N=8192 x 4(of 32) lengthscales, 1 GPU", "wall_ms": t * 1e3,
                      "tflops": m * nsw * N ** 3 / 3 / t / 1e12, "full_C5_8gpu_estimate_s": t * (128 / m) * (32 / nsw) / 8}), flush=True)


if __name__ == "__main__":
    main()
