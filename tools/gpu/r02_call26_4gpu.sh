#!/bin/bash
# Round 2, GPU call 26 (4 GPUs): the partitioned ILMM paths (row-cyclic, distributed storage) at G = 4 -- parity and timings --, then the
# driver's N = 4 bench command.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1"
timeout 500 $TR --master-port 29514 tools/multigpu_ilmm.py 16384 > gpurun_out/r02_ilmm_rowcyclic_4gpu.log 2>&1; echo "ilmm rc=$?"; grep -c " OK" gpurun_out/r02_ilmm_rowcyclic_4gpu.log; grep -c "FAIL" gpurun_out/r02_ilmm_rowcyclic_4gpu.log; grep "^{" gpurun_out/r02_ilmm_rowcyclic_4gpu.log | cut -c1-900
timeout 900 $TR --master-port 29512 bench.py --gpus 4 --steps 3 --warmup 3 > gpurun_out/r02_bench_4gpu.json 2> gpurun_out/r02_c26_bench_4gpu.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r02_c26_bench_4gpu.err; head -c 330 gpurun_out/r02_bench_4gpu.json
