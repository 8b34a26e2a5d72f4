#!/usr/bin/env python
"""Row-cyclic multi-GPU factorisation of one large ILMM factor (option "partition_ilmm"): parity against the
single-GPU schedule, LAPACK and the CPU oracle, then timings of the batch-1 Cholesky with and without the
partition.  Run under torchrun, one rank per GPU."""
import ctypes as C
import json
import os
import sys

import numpy as np
import scipy.linalg as sla
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402
from oracle import lmm_oracle as o  # noqa: E402
from tools.chol_bench import run  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ctx = lmm.Context(local)
    lmm.set_default_context(ctx)
    lmm.dist.init_context_distributed(ctx)
    sizes = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "8192,16384").split(",")]
    ok = True
    # ---- parity 1: the potrf primitive, batch 1
    rng = np.random.default_rng(3)
    N = 3000
    A = rng.standard_normal((N, N))
    A = A @ A.T / N + np.eye(N)
    Lr = sla.cholesky(A, lower=True)
    for part, ob in ((1, 0), (1, 1), (1, 3), (1, 5)):
        ctx.set_option("partition_ilmm", part)
        ctx.set_option("outer_block", ob)
        L, logdet, info = lmm.potrf_batched(A)
        err = float(np.max(np.abs(L[0] - Lr)))
        good = info[0] == 0 and err < 1e-11 and abs(logdet[0] - 2 * np.sum(np.log(np.diag(Lr)))) < 1e-8
        ok &= bool(good)
        print(f"rank {rank}: potrf N={N} schedule={part} outer_block={ob} partitioned max|L-L_lapack|={err:.2e} {'OK' if good else 'FAIL'}", flush=True)
    ctx.set_option("outer_block", 0)
    # ---- parity 2: general ILMM logpdf + posterior + marginals + conditioning + gradient with the partitioned factor
    Nn, p, m, Ns = 900, 6, 3, 40
    x = np.sort(rng.uniform(0, 9, Nn))
    xs = rng.uniform(0, 9, Ns)
    H = rng.uniform(0, 1, (p, m))
    fs = [o.GP(o.Kernel(o.SE, 0.9, 1.2), 0.3), o.GP(o.Kernel(o.MATERN32, 1.3, 0.8), -0.2), o.GP(o.Kernel(o.MATERN52, 0.7, 1.5), 0.1)]
    names = {o.SE: lmm.SEKernel, o.MATERN32: lmm.Matern32Kernel, o.MATERN52: lmm.Matern52Kernel}
    gps = [lmm.GP(g.mean_const, (g.kernel.variance * names[g.kernel.kind]()).compose(lmm.ScaleTransform(g.kernel.inv_lengthscale))) for g in fs]
    y = rng.standard_normal(p * Nn)
    f = lmm.ILMM(lmm.independent_mogp(gps), H)
    O = lmm.MOInputIsotopicByOutputs
    res = {}
    for part in (0, 1):
        ctx.set_option("partition_ilmm", part)
        post, lp = lmm.posterior(f(O(x, p), 0.1), y, with_logpdf=True)
        M, V = lmm.mean_and_var(post(O(xs, p), 0.1))
        lpg, g = lmm.logpdf_and_gradient(f(O(x, p), 0.1), y)
        res[part] = (lp, M, V, lpg, g["H"])
        post.f._owner.free()
    ref = o.ilmm_logpdf(fs, H, x, 0.1, y)
    Mr, Vr = o.ilmm_mean_and_var(o.ilmm_posterior(fs, H, x, 0.1, y), H, xs, 0.1)
    e_lp = abs(res[1][0] - ref) / abs(ref)
    e_m = float(np.max(np.abs(res[1][1] - Mr) / (np.abs(Mr) + 1e-9)))
    e_v = float(np.max(np.abs(res[1][2] - Vr) / np.abs(Vr)))
    e_single = max(abs(res[1][0] - res[0][0]) / abs(ref), float(np.max(np.abs(res[1][1] - res[0][1]))), float(np.max(np.abs(res[1][4] - res[0][4]))))
    good = e_lp < 1e-9 and e_m < 1e-8 and e_v < 1e-8 and e_single < 1e-9
    ok &= bool(good)
    print(f"rank {rank}: ILMM partitioned logpdf rel={e_lp:.2e} mean={e_m:.2e} var={e_v:.2e} vs single-GPU schedule={e_single:.2e} "
          f"{'OK' if good else 'FAIL'}", flush=True)
    # ---- timings: batch-1 Cholesky, single-GPU schedule vs row-cyclic partition (max over ranks)
    for Nb in sizes:
        out = {"N": Nb, "ranks": world}
        for part in (0, 1):
            ctx.set_option("partition_ilmm", part)
            dist.barrier()
            ms, _, ld = run(ctx, Nb, 1, reps=3)
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            key = ["single_gpu", "rowcyclic"][part]
            out[key + "_ms"] = round(float(t.item()), 3)
            out[key + "_logdet"] = ld
        out["speedup"] = round(out["single_gpu_ms"] / out["rowcyclic_ms"], 3)
        out["tflops_rowcyclic"] = round(Nb ** 3 / 3.0 / (out["rowcyclic_ms"] * 1e-3) / 1e12, 2)
        if rank == 0:
            print(json.dumps(out), flush=True)
    ctx.set_option("partition_ilmm", 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
