#!/usr/bin/env python
"""Batch-1 / batch-2 Cholesky schedules (ILMM-shaped work): plain, left-looking K-split look-ahead (1),
right-looking look-ahead (2), over outer_block widths.  Prints JSON lines; tuning aid."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402
from tools.chol_bench import run  # noqa: E402

if __name__ == "__main__":
    ctx = lmm.default_context()
    cfgs = sys.argv[1] if len(sys.argv) > 1 else "4096x1,8192x1,16384x1,8192x2"
    for cfg in cfgs.split(","):
        N, batch = [int(v) for v in cfg.split("x")]
        for la, obs in ((0, [0]), (1, [0]), (2, [1, 2, 3, 4, 5])):
            ctx.set_option("lookahead", la)
            for ob in obs:
                ctx.set_option("outer_block", ob)
                ms, _, ld = run(ctx, N, batch, reps=3)
                print(json.dumps({"N": N, "batch": batch, "lookahead": la, "outer_block": ob or "auto", "chol_ms": round(ms, 3),
                                  "tflops": round(batch * N ** 3 / 3.0 / (ms * 1e-3) / 1e12, 2), "logdet0": ld}), flush=True)
