#!/usr/bin/env python
"""Batch-1 Cholesky (right-looking look-ahead schedule) over the panel-kernel variants: diagonal-tile kernel (potrf_impl 0/1)
x direct GEMM kernel (gemm_direct 0/2) x panel_split x pdl.  Prints JSON lines; tuning aid.   usage: bench_panel_variants.py [N,N,...] [batch]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402
from tools.chol_bench import run  # noqa: E402

if __name__ == "__main__":
    ctx = lmm.default_context()
    Ns = [int(v) for v in (sys.argv[1] if len(sys.argv) > 1 else "2048,4096,8192,16384").split(",")]
    batch = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    for N in Ns:
        for potrf_impl, direct, split, pdl in ((0, 0, 0, 0), (0, 2, 0, 0), (1, 0, 0, 0), (1, 2, 0, 0), (1, 2, 1, 0), (1, 2, 0, 1)):
            ctx.set_option("potrf_impl", potrf_impl)
            ctx.set_option("gemm_direct", direct)
            ctx.set_option("panel_split", split)
            ctx.set_option("pdl", pdl)
            ms, _, ld = run(ctx, N, batch, reps=3)
            print(json.dumps({"N": N, "batch": batch, "potrf_impl": potrf_impl, "gemm_direct": direct, "panel_split": split, "pdl": pdl, "chol_ms": round(ms, 3),
                              "tflops": round(batch * N ** 3 / 3.0 / (ms * 1e-3) / 1e12, 2), "logdet0": ld}), flush=True)
