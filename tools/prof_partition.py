import os, sys
import torch, torch.distributed as dist
sys.path.insert(0, "/root/repo")
import lmm_b200 as lmm
from tools.chol_bench import run
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = lmm.Context(local); lmm.set_default_context(ctx); lmm.dist.init_context_distributed(ctx)
ctx.set_option("partition_ilmm", 2)
for N in (16384,):
    run(ctx, N, 1, reps=1)
    ctx.set_option("profile_partition", 1)
    ms, _, _ = run(ctx, N, 1, reps=0)
    ctx.set_option("profile_partition", 0)
    if dist.get_rank() == 0: print("N", N, "ms", ms, flush=True)
dist.barrier(); dist.destroy_process_group()
