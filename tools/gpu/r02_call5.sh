#!/bin/bash
# Round 2, GPU call 5: fused per-column chain kernel -- schedule tests first (fail fast), full suite, batch-1 timings, notebook.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time timeout 600 python -m pytest tests -m gpu -x -q -k "potrf or schedules or streams_blocking or non_pd" ) > gpurun_out/r02_c5_pytest_chain.log 2>&1
tail -15 gpurun_out/r02_c5_pytest_chain.log
if grep -q "passed" gpurun_out/r02_c5_pytest_chain.log && ! grep -q "failed\|error" gpurun_out/r02_c5_pytest_chain.log; then
  ( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_c5_pytest.log 2>&1
  tail -8 gpurun_out/r02_c5_pytest.log
  python tools/bench_panel_variants.py 1024,2048,4096 1 0,1,2,3 > gpurun_out/r02_c5_panel_variants.jsonl 2>&1
  python tools/bench_panel_variants.py 8192 1 0,2,3,4 >> gpurun_out/r02_c5_panel_variants.jsonl 2>&1
  python tools/bench_panel_variants.py 16384 1 0 >> gpurun_out/r02_c5_panel_variants.jsonl 2>&1
  python tools/bench_panel_variants.py 4096,8192 2 0 >> gpurun_out/r02_c5_panel_variants.jsonl 2>&1
  tail -50 gpurun_out/r02_c5_panel_variants.jsonl
  python tools/bench_notebook.py --reps 20 --no-cpu > gpurun_out/r02_c5_notebook.jsonl 2> gpurun_out/r02_c5_notebook.err; tail -3 gpurun_out/r02_c5_notebook.err
  head -c 600 gpurun_out/r02_c5_notebook.jsonl
fi
