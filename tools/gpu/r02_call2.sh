#!/bin/bash
# Round 2, GPU call 2: persistent solve sweeps + block-Cholesky update under the full parity suite; stage timings of the
# solves (old vs new, 8 and 64 latents); notebook shape again; ncu --set full of the GEMM kernels and the sweep kernels;
# launch list of the notebook-shape eval.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c2_pytest.log 2>&1
tail -8 gpurun_out/r02_c2_pytest.log
for impl in 0 1; do
  python tools/ncu_target.py eval --m 8 --solve-impl $impl --passes 3 2>&1 | tail -1 | sed "s/^/solve_impl=$impl /" >> gpurun_out/r02_c2_solves.log
  python tools/ncu_target.py eval --m 32 --solve-impl $impl --passes 3 2>&1 | tail -1 | sed "s/^/solve_impl=$impl /" >> gpurun_out/r02_c2_solves.log
done
cat gpurun_out/r02_c2_solves.log
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02_c2_bench.json 2> gpurun_out/r02_c2_bench.err; tail -c 400 gpurun_out/r02_c2_bench.err
python tools/bench_notebook.py --reps 20 --no-cpu > gpurun_out/r02_c2_notebook.jsonl 2> gpurun_out/r02_c2_notebook.err; tail -3 gpurun_out/r02_c2_notebook.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
T="python tools/ncu_target.py chol --N 16384 --batch 8"
$T > gpurun_out/r02_c2_ncu_plain.log 2>&1 && {
  $NCU -k 'regex:gemm_tile_kernel_v2<\(int\)0' -s 190 -c 1 -o gpurun_out/r02_ncu_gemm_update_wide -f $T > gpurun_out/r02_c2_ncu1.log 2>&1
  $NCU -k 'regex:gemm_tile_kernel_v2<\(int\)1' -s 191 -c 1 -o gpurun_out/r02_ncu_gemm_trsm -f $T > gpurun_out/r02_c2_ncu2.log 2>&1
}
T2="python tools/ncu_target.py eval --m 8 --passes 2"
$T2 > gpurun_out/r02_c2_ncu_plain2.log 2>&1 && {
  $NCU -k 'regex:fwd_sweep_kernel' -s 1 -c 1 -o gpurun_out/r02_ncu_fwd_sweep -f $T2 > gpurun_out/r02_c2_ncu3.log 2>&1
  $NCU -k 'regex:bwd_sweep_kernel' -s 1 -c 1 -o gpurun_out/r02_ncu_bwd_sweep -f $T2 > gpurun_out/r02_c2_ncu4.log 2>&1
  $NCU -k 'regex:project_kernel' -s 1 -c 1 -o gpurun_out/r02_ncu_project_old -f $T2 > gpurun_out/r02_c2_ncu5.log 2>&1
}
T3="python tools/ncu_target.py notebook"
$T3 > gpurun_out/r02_c2_ncu_plain3.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches_notebook_before.csv $T3 > gpurun_out/r02_c2_ncu6.log 2>&1
ls -la gpurun_out | grep r02_ | tail -30
