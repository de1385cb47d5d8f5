"""Workload for ncu captures: one factorisation + posterior at N (default 16384)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import stheno_jl_b200 as sb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
ns = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
rng = np.random.default_rng(123456)
x = rng.uniform(0, n / 32, n); xs = rng.uniform(0, n / 32, ns)
y = np.sin(x) + 0.3 * rng.standard_normal(n)
f = sb.gppp(lambda GP: dict(f=GP(sb.SEKernel())))
fx = f(sb.GPPPInput("f", x), 0.1)
print(sb.logpdf(fx, y))
post = sb.posterior(fx, y)
m, v = sb.mean_and_var(post, sb.GPPPInput("f", xs))
print(m[:3], v[:3])
