#!/bin/bash
# Round 2, GPU call 23 (1 GPU): the bench default is now 7 planes of 8 bits: whole parity suite with it forced on, the driver's bench command,
# ncu --set full of one wide update of that instantiation.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time LMM_OZAKI=7 LMM_OZAKI_BITS=8 timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c23_pytest_ozaki7x8.log 2>&1; echo "pytest LMM_OZAKI=7 LMM_OZAKI_BITS=8 rc=$?"; tail -6 gpurun_out/r02_c23_pytest_ozaki7x8.log | cut -c1-300
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_ozaki.json 2> gpurun_out/r02_c23_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_c23_bench.err; head -c 300 gpurun_out/r02_bench_1gpu_ozaki.json; echo
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ozaki_update -s 92 -c 1 -o gpurun_out/r02_ncu_ozaki_7x8 -f python tools/ncu_target.py chol --ozaki 7 --ozaki-bits 8 > gpurun_out/r02_c23_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r02_c23_ncu.log
