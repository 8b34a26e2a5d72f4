import sys, time, json
sys.path.insert(0, '/root/repo')
import numpy as np
import lmm_b200 as lmm
from bench import workload
p, m, N, Ns = 64, 64, 16384, 1024
x, U, S, inv_ls, y, s2 = workload(p, m, N)
f = lmm.ILMM(lmm.independent_mogp([lmm.GP(lmm.SEKernel().compose(lmm.ScaleTransform(float(s)))) for s in inv_ls]), lmm.Orthogonal(U, S))
fx = f(lmm.MOInputIsotopicByOutputs(x, p), s2)
post, lp = lmm.posterior(fx, y, with_logpdf=True)
ctx = lmm.default_context()
for Ns in (256, 1024):
    xs = np.random.default_rng(5).uniform(0, N / 100.0, Ns)
    for rep in range(3):
        t0 = time.perf_counter()
        M, V = lmm.mean_and_var(post(lmm.MOInputIsotopicByOutputs(xs, p), s2))
        dt = time.perf_counter() - t0
    print(json.dumps({"config": f"C4 posterior marginals at N*={Ns} (64 latents, N=16384)", "wall_ms": dt * 1e3, "device_ms": float(ctx.last_timings()[0]),
                      "tflops": m * N * N * Ns / dt / 1e12, "post_device_GB": post.f.fs[0]._owner.device_bytes() / 1e9,
                      "mean_range": [float(M.min()), float(M.max())], "var_range": [float(V.min()), float(V.max())]}), flush=True)
# size-independent property: posterior mean at training inputs ~ y_proj - ΣT α on a subset of inputs
