"""Time the triangular-sweep variants (solve.cu) on one factor: python tools/solve_bench.py [N]"""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stheno_jl_b200 as sb
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
x, y, xs = bench.make_inputs(n, 64)
ctx = sb.default_context()
ctx.set_option("trailing", 1)
f = sb.gppp(lambda GP: dict(f=GP(sb.SEKernel())))
fx = f(sb.GPPPInput("f", x), 0.1)
lp0 = sb.logpdf(fx, y)
for var, name in [(-1, "legacy (2 launches/block)"), (0, "A: diag CTA"), (2, "A + L2 prefetch"), (3, "B: owner diag"), (1, "B + L2 prefetch")]:
    ctx.set_option("sweep_variant", var)
    ts = []
    for rep in range(4):
        ctx.timings(reset=True)
        lp = sb.logpdf(fx, y + rep * 1e-3)   # new delta each time: no cache
        ts.append(ctx.timings()["solve_ms"])
    lp = sb.logpdf(fx, y)
    print(f"{name:28s} forward sweep + colsumsq: min {min(ts):7.3f} ms  (all {['%.2f' % t for t in ts]})  logpdf rel diff {abs(lp - lp0) / abs(lp0):.1e}")
