#!/bin/bash
# Round 2, GPU call 11 (1 GPU): first run of the integer-slice (Ozaki) trailing update: parity tests, then a reduced-batch timing.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ozaki_gpu.py -m gpu -x -q > gpurun_out/r02_c11_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r02_c11_pytest.log
for oz in 0 8 7; do
  timeout 300 python bench.py --m 8 --steps 2 --warmup 1 --no-cpu-baseline --no-extras --ozaki $oz > gpurun_out/r02_c11_bench_m8_oz$oz.json 2> gpurun_out/r02_c11_bench_m8_oz$oz.err; echo "bench oz=$oz rc=$?"
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r02_c11_bench_m8_oz$oz.json').read().strip().splitlines()[-1])
    print($oz, d['ms_per_step'], d['stage_ms_per_step'], d.get('parity_check',{}).get('rel_err'))
except Exception as e:
    print('no line', e); print(open('gpurun_out/r02_c11_bench_m8_oz$oz.err').read()[-1500:])
P
done
