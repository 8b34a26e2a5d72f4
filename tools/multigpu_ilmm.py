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


def _maxed(torch, dist, v):
    t = torch.tensor([v], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def ilmm_logpdf_record(lmm, ctx, dist, torch, p, m, N, reps=2):
    """logpdf of a general ILMM (projected form, joint dimension m*N) through the public API on identical inputs on every rank:
    replicated (0), row-cyclic with every rank holding the whole matrix (1), row-cyclic with distributed storage (2).
    Times are CUDA-event stage times of the library (max over ranks): factorisation, and the whole evaluation."""
    rng = np.random.default_rng(5)
    x = np.sort(rng.uniform(0, N / 100.0, N))
    H = rng.uniform(0, 1, (p, m))
    y = rng.standard_normal(p * N)
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP((0.8 + 0.1 * a) * lmm.SEKernel().compose(lmm.ScaleTransform(1.0 + 0.2 * a))) for a in range(m)]), H)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    out = {"config": f"ILMM p={p} m={m} N={N}: projected joint matrix {m * N} x {m * N}", "ranks": dist.get_world_size()}
    for part, key in ((0, "replicated"), (1, "rowcyclic"), (2, "distributed")):
        ctx.set_option("partition_ilmm", part)
        best = (1e30, 1e30)
        for _ in range(reps + 1):
            dist.barrier()
            lp = lmm.logpdf(fx, y)
            tm = ctx.last_timings()
            best = min(best, (_maxed(torch, dist, tm[2]), _maxed(torch, dist, tm[0])))
        out[key + "_factor_ms"], out[key + "_eval_ms"], out[key + "_logpdf"] = best[0], best[1], lp
        if part == 2:
            out["bytes_matrix_per_rank"], out["bytes_workspace_per_rank"], out["bytes_matrix_total"] = tm[4], tm[5], tm[7]
    ctx.set_option("partition_ilmm", 0)
    out["speedup_rowcyclic"] = out["replicated_factor_ms"] / out["rowcyclic_factor_ms"]
    out["speedup_distributed"] = out["replicated_factor_ms"] / out["distributed_factor_ms"]
    out["logpdf_rel_diff"] = max(abs(out[k + "_logpdf"] - out["replicated_logpdf"]) for k in ("rowcyclic", "distributed")) / abs(out["replicated_logpdf"])
    out["tflops_distributed"] = (m * N) ** 3 / 3.0 / (out["distributed_factor_ms"] * 1e-3) / 1e12
    return out


def big_joint_record(lmm, ctx, dist, torch, N, p=8):
    """A joint matrix LARGER than one GPU's memory, factored with distributed storage: ILMM with an orthogonal p x p mixing
    matrix H = U sqrt(S) (dense form, pN x pN), whose logpdf must equal the OILMM's (test/oilmm.jl:10-14's identity), which the
    latent-sharded path evaluates from p factorizations of N x N."""
    from oracle import lmm_oracle as o

    rng = np.random.default_rng(11)
    x = np.sort(rng.uniform(0, N / 100.0, N))
    U, S = o.orthogonal_from_seed(p, p, seed=2)
    y = rng.standard_normal(p * N)
    gps = [lmm.GP((0.8 + 0.05 * a) * lmm.SEKernel().compose(lmm.ScaleTransform(1.0 + 0.1 * a))) for a in range(p)]
    O = lmm.MOInputIsotopicByOutputs
    ctx.set_option("partition_ilmm", 0)
    lp_oilmm = lmm.logpdf(lmm.ILMM(lmm.independent_mogp(gps), lmm.Orthogonal(U, S))(O(x, p), 0.1), y)
    ctx.set_option("partition_ilmm", 2)
    lmm.set_ilmm_form(1)
    dist.barrier()
    lp = lmm.logpdf(lmm.ILMM(lmm.independent_mogp(gps), U * np.sqrt(S)[None, :])(O(x, p), 0.1), y)
    tm = ctx.last_timings()
    lmm.set_ilmm_form(0)
    ctx.set_option("partition_ilmm", 0)
    free, total = torch.cuda.mem_get_info()
    rel = abs(lp - lp_oilmm) / abs(lp_oilmm)
    dim = p * N
    fac = _maxed(torch, dist, tm[2])
    return {"config": f"ILMM dense form p=m={p} N={N}: joint matrix {dim} x {dim}", "ranks": dist.get_world_size(),
            "bytes_matrix_total": tm[7], "gpu_memory_bytes": total, "fits_one_gpu": bool(tm[7] + 2 * 8 * dim < total),
            "bytes_matrix_per_rank": tm[4], "bytes_workspace_per_rank": tm[5], "assemble_ms": _maxed(torch, dist, tm[1]), "factor_ms": fac,
            "eval_ms": _maxed(torch, dist, tm[0]), "tflops_total": dim ** 3 / 3.0 / (fac * 1e-3) / 1e12, "logpdf": lp, "logpdf_oilmm_identity": lp_oilmm,
            "rel_diff": rel, "ok": rel < 1e-9}


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
    # ---- parity 3: distributed storage (partition_ilmm = 2): logpdf of both ILMM forms against the replicated path and the oracle,
    # over block widths / chain variants / a ragged size
    for (Nd, pd, md) in ((900, 6, 3), (1111, 5, 4)):
        xd = np.sort(rng.uniform(0, 9, Nd))
        Hd = rng.uniform(0, 1, (pd, md))
        fsd = [o.GP(o.Kernel([o.SE, o.MATERN32, o.MATERN52, o.SE][a], 0.8 + 0.2 * a, 1.0 + 0.1 * a), 0.1 * a) for a in range(md)]
        gpsd = [lmm.GP(g.mean_const, (g.kernel.variance * names[g.kernel.kind]()).compose(lmm.ScaleTransform(g.kernel.inv_lengthscale))) for g in fsd]
        yd = rng.standard_normal(pd * Nd)
        fd = lmm.ILMM(lmm.independent_mogp(gpsd), Hd)
        refd = o.ilmm_logpdf(fsd, Hd, xd, 0.1, yd)
        for form in (0, 1):
            lmm.set_ilmm_form(form)
            ctx.set_option("partition_ilmm", 0)
            lp0 = lmm.logpdf(fd(O(xd, pd), 0.1), yd)
            for ob, cf in ((0, 1), (1, 1), (3, 0), (5, 2)):
                ctx.set_option("partition_ilmm", 2)
                ctx.set_option("outer_block", ob)
                ctx.set_option("chain_fused", cf)
                lp2 = lmm.logpdf(fd(O(xd, pd), 0.1), yd)
                tm = ctx.last_timings()
                e0, er = abs(lp2 - lp0) / abs(lp0), abs(lp2 - refd) / abs(refd)
                good = e0 < 1e-11 and er < 1e-9 and tm[4] > 0 and tm[4] < 0.75 * tm[7]
                ok &= bool(good)
                print(f"rank {rank}: ILMM distributed storage N={Nd} p={pd} m={md} form={form} outer_block={ob} chain_fused={cf}: vs replicated {e0:.2e} "
                      f"vs oracle {er:.2e} own/full bytes {tm[4] / tm[7]:.3f} {'OK' if good else 'FAIL'}", flush=True)
            ctx.set_option("outer_block", 0)
            ctx.set_option("chain_fused", 1)
        lmm.set_ilmm_form(0)
    ctx.set_option("partition_ilmm", 0)
    # ---- ILMM logpdf at a joint dimension of 16384 (p = 8, m = 4, N = 4096: BASELINE config 2 at twice the N): replicated, row-cyclic,
    # row-cyclic with distributed storage
    if os.environ.get("LMM_ILMM_TIMING", "1") != "0":
        out = ilmm_logpdf_record(lmm, ctx, dist, torch, 8, 4, 4096)
        if rank == 0:
            print(json.dumps(out), flush=True)
    big = int(os.environ.get("LMM_ILMM_BIG", "0"))
    if big:  # a joint matrix that does not fit ONE GPU: p = m = 8 orthogonal H, so the ILMM logpdf must equal the sharded OILMM logpdf
        out = big_joint_record(lmm, ctx, dist, torch, big)
        ok &= bool(out["ok"])
        if rank == 0:
            print(json.dumps(out), flush=True)
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
