#!/bin/bash
# Round 2, GPU call 17 (1 GPU): final single-CTA Ozaki kernel: tests, batch-1 experiment, C4 bench with --ozaki 8, ncu of one wide update.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ozaki_gpu.py -m gpu -q > gpurun_out/r02_c17_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_c17_pytest.log
python - <<'P'
import sys
sys.path.insert(0, '.')
import lmm_b200 as lmm
from tools.chol_bench import run
ctx = lmm.default_context()
for N in (8192, 16384):
    for oz, snt in ((0, 0), (8, 1), (7, 1)):
        ctx.set_option("ozaki", oz); ctx.set_option("ozaki_single_nt", snt)
        for batch in (1, 2):
            ms, _, ld = run(ctx, N, batch, reps=2)
            print(f"N={N} batch={batch} ozaki={oz} single_nt={snt}: cholesky {ms:.2f} ms  {batch*N**3/3/(ms*1e-3)/1e12:.1f} TFLOP/s-eq logdet0 {ld:.6f}", flush=True)
P
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_update -s 44 -c 1 -o gpurun_out/r02_ncu_ozaki_final -f python tools/ncu_target.py chol --ozaki 8 > gpurun_out/r02_c17_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 python bench.py --steps 3 --warmup 1 --no-cpu-baseline --no-extras --ozaki 8 > gpurun_out/r02_c17_bench_oz8.json 2> gpurun_out/r02_c17_bench_oz8.err; echo "bench rc=$?"
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_c17_bench_oz8.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms_per_step'], d['clocks'])
P
