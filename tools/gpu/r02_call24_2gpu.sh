#!/bin/bash
# Round 2, GPU call 24 (2 GPUs): the driver's N = 2 bench command with the final defaults (7 planes of 8 bits).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 1200 $TR --master-port 29512 bench.py --gpus 2 > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_c24_bench_2gpu.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r02_c24_bench_2gpu.err; head -c 400 gpurun_out/r02_bench_2gpu.json
