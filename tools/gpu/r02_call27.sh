#!/bin/bash
# Round 2, GPU call 27 (1 GPU): final code, the driver's sequence: parity suite, smoke, bench (1 GPU), reference arm.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -x -q -m gpu ) > gpurun_out/r02_c27_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r02_c27_pytest.log | cut -c1-200
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c27_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_c27_smoke.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_ozaki.json 2> gpurun_out/r02_c27_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_c27_bench.err; head -c 330 gpurun_out/r02_bench_1gpu_ozaki.json; echo
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r02_ref_1gpu.json 2> gpurun_out/r02_c27_ref.err; echo "ref rc=$?"; tail -c 400 gpurun_out/r02_ref_1gpu.json
