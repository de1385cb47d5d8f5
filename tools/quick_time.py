"""Quick phase timing of the headline pipeline at a few N (development aid, not the bench)."""
import json, sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stheno_jl_b200 as sb

def run(n, ns):
    rng = np.random.default_rng(123456)
    x = rng.uniform(0, n / 32, n); xs = rng.uniform(0, n / 32, ns)
    y = np.sin(x) + 0.3 * rng.standard_normal(n)
    f = sb.gppp(lambda GP: dict(f=GP(sb.SEKernel())))
    ctx = sb.default_context(); ctx.timings(reset=True)
    t0 = time.perf_counter()
    fx = f(sb.GPPPInput("f", x), 0.1)
    lp = sb.logpdf(fx, y)
    post = sb.posterior(fx, y)
    m, v = sb.mean_and_var(post, sb.GPPPInput("f", xs))
    wall = time.perf_counter() - t0
    t = ctx.timings()
    t.update(n=n, ns=ns, wall_s=wall, logpdf=lp, pts_per_s=n / wall,
             trailing_tflops=t["trailing_flops"] / (t["trailing_kernel_ms"] * 1e-3) / 1e12 if t["trailing_kernel_ms"] else 0)
    print(json.dumps(t)); sys.stdout.flush()
    del post, fx

for n, ns in [(4096, 512), (16384, 1024), (32768, 2048), (65536, 4096)]:
    if len(sys.argv) > 1 and n > int(sys.argv[1]): break
    run(n, ns)
    run(n, ns)
