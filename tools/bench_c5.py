#!/usr/bin/env python
"""BASELINE config 5: OILMM p=256, m=128, N=8192, batched hyper-parameter sweep (32 lengthscale
settings per call) with the (sweep x latent) grid of 4096 independent factorizations block-sharded
over the ranks (torchrun, one rank per GPU) and streamed through a fixed arena per GPU.
Prints one JSON line on rank 0.  `--slice k` runs m/k latents and 32/k... for quick checks."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--p", type=int, default=256)
    ap.add_argument("--m", type=int, default=128)
    ap.add_argument("--N", type=int, default=8192)
    ap.add_argument("--sweep", type=int, default=32)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ctx = lmm.Context(local)
    lmm.set_default_context(ctx)
    if world > 1:
        lmm.dist.init_context_distributed(ctx)
    p, m, N = args.p, args.m, args.N
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(0, N / 100.0, N))
    U, S, _ = np.linalg.svd(np.random.default_rng(1).uniform(0, 1, (p, m)), full_matrices=False)
    f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.SEKernel()) for _ in range(m)]), lmm.Orthogonal(U, S))
    y = rng.standard_normal(p * N)
    fx = f(lmm.MOInputIsotopicByOutputs(x, p), 0.1)
    scales = np.geomspace(0.25, 4.0, args.sweep)
    out = lmm.logpdf_sweep(fx, y, scales)  # warm-up
    times = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = lmm.logpdf_sweep(fx, y, scales)
        if world > 1:
            dist.barrier()
        times.append(time.perf_counter() - t0)
    t = torch.tensor([min(times)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if int(os.environ.get("RANK", "0")) == 0:
        sec = float(t.item())
        flops = m * args.sweep * N ** 3 / 3.0
        print(json.dumps({"config": f"C5: OILMM p={p} m={m} N={N} x {args.sweep} lengthscales", "n_gpus": world, "seconds_per_call": sec,
                          "factorizations": m * args.sweep, "tflops_total": flops / sec / 1e12, "tflops_per_gpu": flops / sec / 1e12 / world,
                          "logpdf_min": float(np.min(out)), "logpdf_max": float(np.max(out)), "argmax_scale": float(scales[int(np.argmax(out))])}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
