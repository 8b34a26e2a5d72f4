#!/bin/bash
# Round 2, GPU call 10 (2 GPUs): int8 tcgen05 microbenchmark (GPU 0), row-cyclic ILMM with distributed storage (parity 3 + timing).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 120 tools/microbench/i8_mma > gpurun_out/r02_i8_mma.jsonl 2>&1; echo "i8_mma rc=$?"; cat gpurun_out/r02_i8_mma.jsonl
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 $TR --master-port 29514 tools/multigpu_ilmm.py 16384 > gpurun_out/r02_ilmm_dist_2gpu.log 2>&1; echo "ilmm rc=$?"; grep -v "^rank 1" gpurun_out/r02_ilmm_dist_2gpu.log | tail -40
