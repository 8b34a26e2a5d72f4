#!/bin/bash
# Round 2, GPU call 9 (1 GPU): final code -- full parity suite, the driver's bench command, batch-1 timings with the fused tail,
# ncu launch list (time + DRAM bytes) of the bench command for roofline.traffic.
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_c9_pytest.log 2>&1
tail -6 gpurun_out/r02_c9_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_c9_smoke.log 2>&1; tail -2 gpurun_out/r02_c9_smoke.log
python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_c9_bench.err; tail -c 300 gpurun_out/r02_c9_bench.err; head -c 400 gpurun_out/r02_bench_1gpu.json
python tools/bench_batch1.py 1024x1,2048x1,4096x1,6144x1,8192x1,12288x1,16384x1,8192x2 1 0 74 > gpurun_out/r02_c9_batch1.jsonl 2>&1; cat gpurun_out/r02_c9_batch1.jsonl
python tools/bench_configs.py > gpurun_out/r02_c9_configs.jsonl 2> gpurun_out/r02_c9_configs.err; cut -c1-330 gpurun_out/r02_c9_configs.jsonl
python tools/c4_marginals.py > gpurun_out/r02_c9_c4_marginals.log 2>&1; tail -3 gpurun_out/r02_c9_c4_marginals.log
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras --streams 1"
$B > gpurun_out/r02_c9_ncu_plain.log 2>&1 &&
  ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/r02_launches_c4.csv $B > gpurun_out/r02_c9_ncu.log 2>&1
wc -l gpurun_out/r02_launches_c4.csv
