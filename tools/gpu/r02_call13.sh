#!/bin/bash
# Round 2, GPU call 13 (1 GPU): Ozaki update with the unrolled issue loop: parity tests, timing at 8 and 64 latents, ncu of one wide update.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ozaki_gpu.py -m gpu -q > gpurun_out/r02_c13_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_c13_pytest.log
for oz in 0 8 7; do
  timeout 300 python tools/ncu_target.py chol --ozaki $oz 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_update -s 22 -c 1 -o gpurun_out/r02_ncu_ozaki_update_v2 -f python tools/ncu_target.py chol --ozaki 8 > gpurun_out/r02_c13_ncu.log 2>&1; echo "ncu rc=$?"
for oz in 8; do
  timeout 600 python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-extras --ozaki $oz > gpurun_out/r02_c13_bench_oz$oz.json 2> gpurun_out/r02_c13_bench_oz$oz.err; echo "bench oz=$oz rc=$?"
  python - <<P
import json
try:
    d=json.loads(open('gpurun_out/r02_c13_bench_oz$oz.json').read().strip().splitlines()[-1])
    print($oz, d['ms_per_step'], d['stage_ms_per_step'], d.get('parity_check'))
except Exception as e:
    print('no line', e); print(open('gpurun_out/r02_c13_bench_oz$oz.err').read()[-1500:])
P
done
