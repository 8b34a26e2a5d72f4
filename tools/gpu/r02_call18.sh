#!/bin/bash
# Round 2, GPU call 18 (1 GPU): full parity suite (library default = DMMA), smoke, the driver's bench command (timed path = int8 update),
# ncu launch list of one step.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -q ) > gpurun_out/r02_c18_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r02_c18_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c18_smoke.log 2>&1; tail -1 gpurun_out/r02_c18_smoke.log
timeout 1500 python bench.py > gpurun_out/r02_bench_1gpu_ozaki.json 2> gpurun_out/r02_c18_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_c18_bench.err; head -c 600 gpurun_out/r02_bench_1gpu_ozaki.json
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --streams 1"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4_ozaki.csv $B > gpurun_out/r02_c18_ncu.log 2>&1; echo "ncu rc=$?"
wc -l gpurun_out/r02_launches_c4_ozaki.csv
