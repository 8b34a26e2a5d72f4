#!/bin/bash
# Round 2, GPU call 21 (1 GPU): final code.  Full parity suite on the library default (DMMA) and once more with the int8 update forced on
# (LMM_OZAKI=8), smoke, the driver's bench command, ncu launch list of one eval of the default bench path, batch-1 table.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c21_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_c21_pytest.log
( time LMM_OZAKI=8 timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c21_pytest_ozaki8.log 2>&1; echo "pytest LMM_OZAKI=8 rc=$?"; tail -6 gpurun_out/r02_c21_pytest_ozaki8.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c21_smoke.log 2>&1; tail -1 gpurun_out/r02_c21_smoke.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_ozaki.json 2> gpurun_out/r02_c21_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_c21_bench.err; head -c 300 gpurun_out/r02_bench_1gpu_ozaki.json; echo
python tools/bench_batch1.py 4096x1,8192x1,16384x1,8192x2,16384x2 1 0 74 > gpurun_out/r02_c21_batch1.jsonl 2>&1; cat gpurun_out/r02_c21_batch1.jsonl
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --no-dmma --streams 1"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4_ozaki.csv $B > gpurun_out/r02_c21_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_launches_c4_ozaki.csv
