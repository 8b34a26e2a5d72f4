#!/bin/bash
# Round 2, GPU call 19 (2 GPUs): multi-rank parity of every sharded entry point, row-cyclic ILMM incl. distributed storage and a joint
# matrix LARGER than one GPU's memory (229376^2, 210 GB packed), the driver's N = 2 bench command, the reference arm under torchrun.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29511 tools/multigpu_check.py > gpurun_out/r02_multigpu_check_2gpu.log 2>&1; echo "check rc=$?"; tail -3 gpurun_out/r02_multigpu_check_2gpu.log
LMM_ILMM_BIG=28672 timeout 900 $TR --master-port 29514 tools/multigpu_ilmm.py 16384 > gpurun_out/r02_ilmm_rowcyclic_2gpu.log 2>&1; echo "ilmm rc=$?"; grep -v "^rank 1" gpurun_out/r02_ilmm_rowcyclic_2gpu.log | grep "^{" | cut -c1-900
timeout 1200 $TR --master-port 29512 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_c19_bench_2gpu.err; echo "bench rc=$?"; tail -c 600 gpurun_out/r02_c19_bench_2gpu.err; tail -c 2500 gpurun_out/r02_bench_2gpu.json
timeout 600 $TR --master-port 29513 bench.py --impl reference --gpus 2 --steps 1 --warmup 1 > gpurun_out/r02_ref_2gpu.json 2> gpurun_out/r02_c19_ref_2gpu.err; echo "ref rc=$?"; tail -c 500 gpurun_out/r02_ref_2gpu.json
