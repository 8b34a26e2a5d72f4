#!/bin/bash
# Round 2, GPU call 16 (1 GPU): CTA-pair (cta_group::2) Ozaki kernel: parity tests, timing single vs pair, ncu of one wide update.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ozaki_gpu.py -m gpu -q -x > gpurun_out/r02_c16_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_c16_pytest.log
python - <<'P'
import sys
sys.path.insert(0, '.')
import lmm_b200 as lmm
from tools.chol_bench import run
ctx = lmm.default_context()
for oz, pair in ((8, 0), (8, 1), (7, 1)):
    ctx.set_option("ozaki", oz); ctx.set_option("ozaki_pair", pair)
    ms, _, ld = run(ctx, 16384, 16, reps=2)
    print(f"ozaki={oz} pair={pair}: batch 16 N=16384 cholesky {ms:.2f} ms  {16*16384**3/3/(ms*1e-3)/1e12:.1f} TFLOP/s-equivalent logdet0 {ld:.6f}", flush=True)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_update_pair -s 44 -c 1 -o gpurun_out/r02_ncu_ozaki_pair -f python tools/ncu_target.py chol --ozaki 8 > gpurun_out/r02_c16_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_c16_ncu.log
