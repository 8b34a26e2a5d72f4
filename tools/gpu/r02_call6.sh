#!/bin/bash
# Round 2, GPU call 6: heterotopic additions + tuned fused-chain rule under the full suite; batch-1 timings; notebook.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c6_pytest.log 2>&1
tail -8 gpurun_out/r02_c6_pytest.log
python tools/bench_batch1.py 1024x1,2048x1,4096x1,8192x1,16384x1,8192x2 1 0 74 > gpurun_out/r02_c6_batch1.jsonl 2>&1
cat gpurun_out/r02_c6_batch1.jsonl
python tools/bench_notebook.py --reps 30 > gpurun_out/r02_c6_notebook.jsonl 2> gpurun_out/r02_c6_notebook.err; tail -3 gpurun_out/r02_c6_notebook.err
python tools/bench_configs.py > gpurun_out/r02_c6_configs.jsonl 2> gpurun_out/r02_c6_configs.err; tail -3 gpurun_out/r02_c6_configs.err; cat gpurun_out/r02_c6_configs.jsonl
