#!/bin/bash
# Round 2, GPU call 1: parity suite with the new full-shape tests, bench (both arms), the reference's published shape before
# any optimisation, and ncu --set full on the final round-1 kernels.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total,clocks.max.sm --format=csv > gpurun_out/r02_c1_gpu.txt
nproc >> gpurun_out/r02_c1_gpu.txt; free -g >> gpurun_out/r02_c1_gpu.txt
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c1_pytest.log 2>&1
tail -5 gpurun_out/r02_c1_pytest.log
python bench.py --steps 3 --warmup 3 > gpurun_out/r02_c1_bench.json 2> gpurun_out/r02_c1_bench.err; tail -c 600 gpurun_out/r02_c1_bench.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_c1_ref.json 2> gpurun_out/r02_c1_ref.err
python bench.py --impl reference --steps 8 --warmup 2 > gpurun_out/r02_c1_ref_once.json 2>> gpurun_out/r02_c1_ref.err
python tools/bench_notebook.py --reps 20 > gpurun_out/r02_c1_notebook.jsonl 2> gpurun_out/r02_c1_notebook.err; tail -3 gpurun_out/r02_c1_notebook.err
T="python tools/ncu_target.py chol --N 16384 --batch 8"
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
$T > gpurun_out/r02_c1_ncu_plain.log 2>&1 && {
  $NCU -k 'regex:gemm_tile_kernel_v2<0' -s 190 -c 1 -o gpurun_out/r02_ncu_gemm_update_wide -f $T > gpurun_out/r02_c1_ncu1.log 2>&1
  $NCU -k 'regex:gemm_tile_kernel_v2<1' -s 191 -c 1 -o gpurun_out/r02_ncu_gemm_trsm -f $T > gpurun_out/r02_c1_ncu2.log 2>&1
  $NCU -k 'regex:potrf_tile_kernel2' -s 192 -c 1 -o gpurun_out/r02_ncu_potrf_tile -f $T > gpurun_out/r02_c1_ncu3.log 2>&1
  $NCU -k 'regex:kmat_sym_kernel' -s 1 -c 1 -o gpurun_out/r02_ncu_kmat_sym -f $T > gpurun_out/r02_c1_ncu4.log 2>&1
}
ls -la gpurun_out | tail -20
