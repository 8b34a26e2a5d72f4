#!/usr/bin/env python
"""Short, deterministic launch sequences for `ncu --set full` captures of the individual kernels (one CUDA stream, so the
n-th launch of a kernel name is always the same piece of work).

    python tools/ncu_target.py chol   [--N 16384 --batch 8]   kernel-matrix build + batched blocked Cholesky (lmm_potrf_bench), twice
    python tools/ncu_target.py eval   [--N 16384 --m 8 --p 64] one OILMM logpdf + posterior (projection, kmat, Cholesky, solves), twice
    python tools/ncu_target.py notebook                        the reference's published shape (p=600, m=20, N=552), OILMM logpdf x 3

With streams = 1 and outer_block = 8 one `chol` pass launches, in order: kmat_sym_kernel<1>; then per block of 8 tile columns
[wide update gemm_tile_kernel_v2<0,32> (not for the first block)] and per column [narrow in-block update (not the block's
first column), potrf_tile_kernel2, TRSM gemm_tile_kernel_v2<1,32>].  At N = 16384 (128 tile columns): 127 UPDATE launches,
128 potrf, 127 TRSM per pass; the wide update of block 8 (K = 64 k-tiles, 8 x 64 x batch tiles) is UPDATE launch #63 of a
pass, i.e. `-k regex:gemm_tile_kernel_v2<0 -s $((127 + 63)) -c 1` for the second (warm) pass."""
import argparse
import ctypes as C
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402
from lmm_b200._lib import GpDesc, ptr  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("mode", choices=["chol", "eval", "notebook"])
    ap.add_argument("--N", type=int, default=16384)
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--m", type=int, default=8)
    ap.add_argument("--p", type=int, default=64)
    ap.add_argument("--passes", type=int, default=2)
    ap.add_argument("--solve-impl", type=int, default=1)
    ap.add_argument("--ozaki", type=int, default=0, help="int8 digit planes of the optional integer-slice trailing update (0 = DMMA)")
    ap.add_argument("--ozaki-bits", type=int, default=7, help="bits per digit plane (7: radix 128, 8: radix 256)")
    args = ap.parse_args()
    ctx = lmm.default_context()
    ctx.set_option("streams", 1)
    ctx.set_option("solve_impl", args.solve_impl)
    ctx.set_option("ozaki", args.ozaki)
    ctx.set_option("ozaki_bits", args.ozaki_bits)
    rng = np.random.default_rng(0)
    if args.mode == "chol":
        x = np.sort(rng.uniform(0, args.N / 100.0, args.N))
        d = GpDesc(0, 0, 1.0, 1.3, 0.0, None, 1.0)
        logdet = np.zeros(args.batch)
        a, c = C.c_double(), C.c_double()
        for _ in range(args.passes):
            rc = ctx.lib.lmm_potrf_bench(ctx.handle, C.byref(d), ptr(x), args.N, 1, 0.1, args.batch, ptr(logdet), C.byref(a), C.byref(c))
            assert rc == 0, (rc, ctx.error())
            print(f"chol N={args.N} batch={args.batch}: kmat {a.value:.3f} ms, cholesky {c.value:.3f} ms, "
                  f"{args.batch * args.N ** 3 / 3 / (c.value * 1e-3) / 1e12:.2f} TFLOP/s, logdet0 {logdet[0]:.6f}", flush=True)
    elif args.mode == "eval":
        from bench import workload

        x, U, S, inv_ls, y, s2 = workload(args.p, args.m, args.N)
        f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.SEKernel().compose(lmm.ScaleTransform(float(s)))) for s in inv_ls]), lmm.Orthogonal(U, S))
        fx = f(lmm.MOInputIsotopicByOutputs(x, args.p), s2)
        for _ in range(args.passes):
            post, lp = lmm.posterior(fx, y, with_logpdf=True)
            tm = ctx.last_timings()
            print(f"eval p={args.p} m={args.m} N={args.N}: logpdf {lp:.6f}; total {tm[0]:.3f} ms, project {tm[4]:.3f}, kmat {tm[1]:.3f}, "
                  f"cholesky {tm[2]:.3f}, solves {tm[3]:.3f}", flush=True)
            post.f.fs[0]._owner.free()
    else:
        sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
        from bench_notebook import notebook_problem

        p, m, U, S, x, xs, rng = notebook_problem()
        y = rng.standard_normal(p * len(x))
        f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.Matern52Kernel()) for _ in range(m)]), lmm.Orthogonal(U, S))
        fx = f(lmm.MOInputIsotopicByOutputs(x, p), 1e-6)
        for _ in range(3):
            lp = lmm.logpdf(fx, y)
            print(f"notebook shape: logpdf {lp:.6f}, device {ctx.last_timings()[0]:.3f} ms", flush=True)


if __name__ == "__main__":
    main()
