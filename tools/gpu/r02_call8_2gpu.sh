#!/bin/bash
# Round 2, GPU call 8 (2 GPUs): the driver's N>1 bench command with the new untimed records (sharded_vs_unsharded, parity_check,
# ilmm_rowcyclic, c5_sweep), the reference arm under torchrun, multi-rank parity incl. the collective PosDef verdict, row-cyclic ILMM.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 tools/multigpu_check.py > gpurun_out/r02_multigpu_check_2gpu.log 2>&1; tail -4 gpurun_out/r02_multigpu_check_2gpu.log
$TR --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/r02_c8_bench_2gpu.json 2> gpurun_out/r02_c8_bench_2gpu.err; tail -c 1500 gpurun_out/r02_c8_bench_2gpu.err; tail -c 3000 gpurun_out/r02_c8_bench_2gpu.json
$TR --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_c8_ref_2gpu.json 2> gpurun_out/r02_c8_ref_2gpu.err; tail -c 700 gpurun_out/r02_c8_ref_2gpu.json
$TR --master-port 29514 tools/multigpu_ilmm.py 8192,16384 > gpurun_out/r02_ilmm_rowcyclic_2gpu.log 2>&1; tail -8 gpurun_out/r02_ilmm_rowcyclic_2gpu.log
