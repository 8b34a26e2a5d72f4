#!/usr/bin/env python
"""Batch-1 / small-batch Cholesky over the panel-chain variants: fused per-column chain kernel (chain_fused 0/1) x programmatic
dependent launch (pdl 0/1) x block width.  Prints JSON lines; tuning aid.   usage: bench_panel_variants.py [N,N,...] [batch] [outer_blocks]"""
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
    obs = [int(v) for v in (sys.argv[3] if len(sys.argv) > 3 else "0").split(",")]
    for N in Ns:
        for ob in obs:
            for fused, pdl in ((0, 0), (0, 1), (1, 0), (1, 1)):
                ctx.set_option("outer_block", ob)
                ctx.set_option("chain_fused", fused)
                ctx.set_option("pdl", pdl)
                ms, _, ld = run(ctx, N, batch, reps=3)
                print(json.dumps({"N": N, "batch": batch, "outer_block": ob or "auto", "chain_fused": fused, "pdl": pdl, "chol_ms": round(ms, 3),
                                  "tflops": round(batch * N ** 3 / 3.0 / (ms * 1e-3) / 1e12, 2), "logdet0": ld}), flush=True)
