#!/bin/bash
# Round 2, GPU call 20 (1 GPU): block-width / min_k sweep of the int8 path, the 8-latents-per-GPU case (what each rank sees at 8 GPUs).
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
python - <<'P' 2>&1 | tee gpurun_out/r02_c20_sweep.log
import sys
sys.path.insert(0, '.')
import lmm_b200 as lmm
from tools.chol_bench import run
ctx = lmm.default_context()
for batch in (16, 8):
    for oz, ob, mk in ((8, 0, 8), (8, 2, 8), (8, 2, 4), (8, 1, 8), (8, 3, 6), (0, 0, 8)):
        ctx.set_option("ozaki", oz); ctx.set_option("outer_block", ob); ctx.set_option("ozaki_min_k", mk)
        ms, _, ld = run(ctx, 16384, batch, reps=2)
        print(f"batch={batch} ozaki={oz} outer_block={ob} min_k={mk}: cholesky {ms:.2f} ms  {batch*16384**3/3/(ms*1e-3)/1e12:.1f} TFLOP/s-eq logdet0 {ld:.6f}", flush=True)
for streams in (1, 2, 8):
    ctx.set_option("ozaki", 8); ctx.set_option("outer_block", 0); ctx.set_option("ozaki_min_k", 8); ctx.set_option("streams", streams)
    ms, _, ld = run(ctx, 16384, 8, reps=2)
    print(f"batch=8 ozaki=8 streams={streams}: cholesky {ms:.2f} ms", flush=True)
P
for oz in 8 0; do
  timeout 300 python bench.py --m 8 --steps 3 --warmup 2 --no-cpu-baseline --no-extras --ozaki $oz > gpurun_out/r02_c20_bench_m8_oz$oz.json 2> gpurun_out/r02_c20_bench_m8_oz$oz.err; echo "bench m=8 oz=$oz rc=$?"
  python - <<P
import json
d=json.loads(open('gpurun_out/r02_c20_bench_m8_oz$oz.json').read().strip().splitlines()[-1])
print($oz, d['ms_per_step'], d['stage_ms_per_step'], d['roofline'].get('frac'), d['roofline'].get('kernel_share_of_step'))
P
done
