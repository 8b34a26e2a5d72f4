#!/bin/bash
# Round 2, GPU call 12 (1 GPU): Ozaki trailing update with the deeper pass-A ring: parity tests, reduced-batch timing, ncu --set full of
# one wide int8 update (block 8: K = 64 k-tiles).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ozaki_gpu.py -m gpu -q > gpurun_out/r02_c12_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r02_c12_pytest.log
for oz in 0 8 7; do
  timeout 300 python tools/ncu_target.py chol --ozaki $oz 2>&1 | tail -1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_update -s 22 -c 1 -o gpurun_out/r02_ncu_ozaki_update -f python tools/ncu_target.py chol --ozaki 8 > gpurun_out/r02_c12_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 gpurun_out/r02_c12_ncu.log
timeout 600 ncu --set full --clock-control none -k regex:ozaki_slice -s 22 -c 1 -o gpurun_out/r02_ncu_ozaki_slice -f python tools/ncu_target.py chol --ozaki 8 > gpurun_out/r02_c12_ncu2.log 2>&1; echo "ncu rc=$?"
