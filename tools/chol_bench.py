#!/usr/bin/env python
"""Kernel-level benchmark of the batched blocked Cholesky (lmm_potrf_bench): sweeps N, batch,
streams, outer_block and prints TFLOP/s (algorithmic N^3/3 per matrix) -- tuning aid."""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import lmm_b200 as lmm  # noqa: E402
from lmm_b200._lib import GpDesc, ptr  # noqa: E402


def run(ctx, N, batch, reps=2):
    rng = np.random.default_rng(0)
    x = np.sort(rng.uniform(0, N / 100.0, N))
    d = GpDesc(0, 0, 1.0, 1.0, 0.0, None, 1.0)
    logdet = np.zeros(batch)
    a, c = C.c_double(), C.c_double()
    best = 1e30
    for _ in range(reps + 1):
        rc = ctx.lib.lmm_potrf_bench(ctx.handle, C.byref(d), ptr(x), N, 1, 0.01, batch, ptr(logdet), C.byref(a), C.byref(c))
        assert rc == 0, (rc, ctx.error())
        best = min(best, c.value)
    return best, a.value, logdet[0]


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--configs", default="8192x8,16384x8,4096x16,2048x32")
    ap.add_argument("--streams", default="1,2,4,8")
    ap.add_argument("--outer", default="8")
    args = ap.parse_args()
    ctx = lmm.default_context()
    for cfg in args.configs.split(","):
        N, batch = [int(v) for v in cfg.split("x")]
        for ob in [int(v) for v in args.outer.split(",")]:
            ctx.set_option("outer_block", ob)
            for st in [int(v) for v in args.streams.split(",")]:
                ctx.set_option("streams", st)
                ms, ms_k, ld = run(ctx, N, batch)
                tf = batch * N ** 3 / 3.0 / (ms * 1e-3) / 1e12
                print(json.dumps({"N": N, "batch": batch, "outer_block": ob, "streams": st, "chol_ms": round(ms, 3), "kmat_ms": round(ms_k, 3),
                                  "tflops": round(tf, 2), "logdet0": ld}), flush=True)
