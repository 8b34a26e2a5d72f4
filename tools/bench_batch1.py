#!/usr/bin/env python
"""Batch-1 / batch-2 Cholesky schedules (ILMM-shaped work): plain (0), right-looking look-ahead on two streams (1), over
outer_block widths and the small-grid threshold of the latency-optimised GEMM kernel.
Prints JSON lines; tuning aid.   usage: bench_batch1.py [NxB,NxB,...] [lookaheads] [outer_blocks] [gemm_small values]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402
from tools.chol_bench import run  # noqa: E402

if __name__ == "__main__":
    ctx = lmm.default_context()
    arg = lambda i, d: sys.argv[i] if len(sys.argv) > i else d
    cfgs = arg(1, "4096x1,8192x1,16384x1,8192x2")
    las = [int(v) for v in arg(2, "1").split(",")]
    obs = [int(v) for v in arg(3, "0,1,2,3,4").split(",")]
    smalls = [int(v) for v in arg(4, "74").split(",")]
    for cfg in cfgs.split(","):
        N, batch = [int(v) for v in cfg.split("x")]
        for la in las:
            ctx.set_option("lookahead", la)
            for ob in obs:
                ctx.set_option("outer_block", ob)
                for small in smalls:
                    ctx.set_option("gemm_small", small)
                    ms, _, ld = run(ctx, N, batch, reps=3)
                    print(json.dumps({"N": N, "batch": batch, "lookahead": la, "gemm_small": small, "outer_block": ob or "auto",
                                      "chol_ms": round(ms, 3), "tflops": round(batch * N ** 3 / 3.0 / (ms * 1e-3) / 1e12, 2), "logdet0": ld}), flush=True)
