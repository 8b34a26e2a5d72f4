#!/bin/bash
# Round 2, GPU call 28 (1 GPU): the two-issuer int8 kernel: whole parity suite with the path forced on, the driver's bench command,
# ncu --set full of one wide update.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time LMM_OZAKI=7 LMM_OZAKI_BITS=8 timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c28_pytest_ozaki7x8.log 2>&1; echo "pytest LMM_OZAKI=7 LMM_OZAKI_BITS=8 rc=$?"; tail -6 gpurun_out/r02_c28_pytest_ozaki7x8.log | cut -c1-300
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c28_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_c28_smoke.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_ozaki.json 2> gpurun_out/r02_c28_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_c28_bench.err; head -c 330 gpurun_out/r02_bench_1gpu_ozaki.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_update -s 92 -c 1 -o gpurun_out/r02_ncu_ozaki_7x8_v2 -f python tools/ncu_target.py chol --ozaki 7 --ozaki-bits 8 > gpurun_out/r02_c28_ncu.log 2>&1; echo "ncu rc=$?"
