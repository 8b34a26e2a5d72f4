#!/bin/bash
# Round 2, GPU call 7 (1 GPU): full parity suite after the heterotopic test fix.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_c7_pytest.log 2>&1
tail -12 gpurun_out/r02_c7_pytest.log
python tools/bench_batch1.py 8192x1,6144x1,12288x1 1 0 74 > gpurun_out/r02_c7_batch1.jsonl 2>&1; cat gpurun_out/r02_c7_batch1.jsonl
