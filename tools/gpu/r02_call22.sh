#!/bin/bash
# Round 2, GPU call 22 (1 GPU): radix-256 digit planes (ozaki_bits = 8): parity tests, Cholesky timing, C4 bench with 7 planes of 8 bits.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ozaki_gpu.py -m gpu -q > gpurun_out/r02_c22_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_c22_pytest.log | cut -c1-300
python - <<'P' 2>&1 | tee gpurun_out/r02_c22_sweep.log
import sys
sys.path.insert(0, '.')
import lmm_b200 as lmm
from tools.chol_bench import run
ctx = lmm.default_context()
for batch in (16, 8):
    for oz, bits in ((8, 7), (7, 8), (6, 8)):
        ctx.set_option("ozaki", oz); ctx.set_option("ozaki_bits", bits)
        ms, _, ld = run(ctx, 16384, batch, reps=2)
        print(f"batch={batch} planes={oz} bits={bits}: cholesky {ms:.2f} ms  {batch*16384**3/3/(ms*1e-3)/1e12:.1f} TFLOP/s-eq logdet0 {ld:.9f}", flush=True)
P
timeout 900 python bench.py --ozaki 7 --ozaki-bits 8 --no-cpu-baseline > gpurun_out/r02_c22_bench_7x8.json 2> gpurun_out/r02_c22_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_c22_bench.err
python - <<'P'
import json
d=json.loads(open('gpurun_out/r02_c22_bench_7x8.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['stage_ms_per_step'], d['dmma_path']['logpdf_rel_diff'], d['dmma_path']['mean_relnorm'], d['dmma_path']['var_max_rel'], d['roofline']['frac'])
P
