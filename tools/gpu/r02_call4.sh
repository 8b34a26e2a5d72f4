#!/bin/bash
# Round 2, GPU call 4: composite kernels + K-split projection under the full suite; notebook shape; ncu of the projection.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c4_pytest.log 2>&1
tail -8 gpurun_out/r02_c4_pytest.log
python tools/bench_notebook.py --reps 20 --no-cpu > gpurun_out/r02_c4_notebook.jsonl 2> gpurun_out/r02_c4_notebook.err; tail -3 gpurun_out/r02_c4_notebook.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c4_bench.json 2> gpurun_out/r02_c4_bench.err; tail -c 300 gpurun_out/r02_c4_bench.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
T2="python tools/ncu_target.py eval --m 64 --p 64 --N 16384 --passes 1"
T3="python tools/ncu_target.py notebook"
$T3 > gpurun_out/r02_c4_ncu_plain3.log 2>&1 &&
  $NCU -k 'regex:project_dmma_kernel' -s 2 -c 1 -o gpurun_out/r02_ncu_project_dmma_notebook -f $T3 > gpurun_out/r02_c4_ncu1.log 2>&1
$T2 > gpurun_out/r02_c4_ncu_plain2.log 2>&1 &&
  $NCU -k 'regex:project_dmma_kernel' -s 0 -c 1 -o gpurun_out/r02_ncu_project_dmma_c4 -f $T2 > gpurun_out/r02_c4_ncu3.log 2>&1
ls -la gpurun_out | grep r02_c4
