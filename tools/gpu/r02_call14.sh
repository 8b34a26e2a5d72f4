#!/bin/bash
# Round 2, GPU call 14 (1 GPU): the WHOLE GPU parity suite with the integer-slice trailing update switched on (LMM_OZAKI=8), then block-width
# sweep of the C4 Cholesky with it.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time LMM_OZAKI=8 timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c14_pytest_ozaki8.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_c14_pytest_ozaki8.log
python - <<'P'
import sys, ctypes as C, numpy as np
sys.path.insert(0, '.')
import lmm_b200 as lmm
from tools.chol_bench import run
ctx = lmm.default_context()
for oz, ob, mk in ((8, 0, 8), (8, 4, 4), (8, 4, 8), (8, 6, 6), (8, 12, 12), (8, 16, 16), (7, 0, 8)):
    ctx.set_option("ozaki", oz); ctx.set_option("outer_block", ob); ctx.set_option("ozaki_min_k", mk)
    ms, _, ld = run(ctx, 16384, 16, reps=2)
    print(f"ozaki={oz} outer_block={ob} min_k={mk}: batch 16 N=16384 cholesky {ms:.2f} ms  {16*16384**3/3/(ms*1e-3)/1e12:.1f} TFLOP/s-equivalent logdet0 {ld:.6f}", flush=True)
P
