#!/bin/bash
# Round 2, GPU call 3: DMMA projection kernel -- parity subset, notebook shape, C4 stage time, ncu.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q -k "project or oilmm or golden or reference or c4_full or c5 or truth or ilmm" ) > gpurun_out/r02_c3_pytest.log 2>&1
tail -8 gpurun_out/r02_c3_pytest.log
python tools/bench_notebook.py --reps 20 --no-cpu > gpurun_out/r02_c3_notebook.jsonl 2> gpurun_out/r02_c3_notebook.err; tail -3 gpurun_out/r02_c3_notebook.err
for m in 8 32; do python tools/ncu_target.py eval --m $m --passes 3 2>&1 | tail -1 >> gpurun_out/r02_c3_eval.log; done
python tools/ncu_target.py eval --m 64 --p 64 --N 4096 --passes 3 2>&1 | tail -1 >> gpurun_out/r02_c3_eval.log
cat gpurun_out/r02_c3_eval.log
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
T2="python tools/ncu_target.py eval --m 64 --p 64 --N 16384 --passes 1"
T3="python tools/ncu_target.py notebook"
$T3 > gpurun_out/r02_c3_ncu_plain3.log 2>&1 && {
  $NCU -k 'regex:project_dmma_kernel' -s 2 -c 1 -o gpurun_out/r02_ncu_project_dmma_notebook -f $T3 > gpurun_out/r02_c3_ncu1.log 2>&1
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_notebook_after_proj.csv $T3 > gpurun_out/r02_c3_ncu2.log 2>&1
}
$T2 > gpurun_out/r02_c3_ncu_plain2.log 2>&1 &&
  $NCU -k 'regex:project_dmma_kernel' -s 0 -c 1 -o gpurun_out/r02_ncu_project_dmma_c4 -f $T2 > gpurun_out/r02_c3_ncu3.log 2>&1
ls -la gpurun_out | grep r02_c3
