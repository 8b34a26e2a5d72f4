#!/bin/bash
# Round 2, GPU call 25 (1 GPU): the int8 path in the prediction sweep (X <- X L^{-T}): int8 tests, C4 marginals with / without, C3 with,
# then the whole parity suite with the option forced on.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ozaki_gpu.py -m gpu -q -x > gpurun_out/r02_c25_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r02_c25_pytest.log | cut -c1-300
for oz in 0 7; do
  LMM_OZAKI=$oz LMM_OZAKI_BITS=8 timeout 600 python tools/c4_marginals.py > gpurun_out/r02_c25_c4_marginals_oz$oz.log 2>&1; echo "c4 marginals oz=$oz rc=$?"; tail -3 gpurun_out/r02_c25_c4_marginals_oz$oz.log | cut -c1-400
done
LMM_OZAKI=7 LMM_OZAKI_BITS=8 timeout 600 python tools/bench_configs.py > gpurun_out/r02_c25_configs_oz7x8.jsonl 2> gpurun_out/r02_c25_configs.err; cut -c1-420 gpurun_out/r02_c25_configs_oz7x8.jsonl
( time LMM_OZAKI=7 LMM_OZAKI_BITS=8 timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c25_pytest_ozaki7x8.log 2>&1; echo "pytest LMM_OZAKI=7 LMM_OZAKI_BITS=8 rc=$?"; tail -6 gpurun_out/r02_c25_pytest_ozaki7x8.log | cut -c1-300
