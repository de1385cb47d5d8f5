import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, scipy.linalg as sla
import stheno_jl_b200 as sb
n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
rng = np.random.default_rng(123456)
x = rng.uniform(0, n / 32, n)
f = sb.gppp(lambda GP: dict(f=GP(sb.SEKernel())))
K = sb.cov(f, sb.GPPPInput("f", x)); K[np.diag_indices(n)] += 0.1
Lref = sla.cholesky(K, lower=True)
for trial in range(30):
    fx = f(sb.GPPPInput("f", x), 0.1)
    try:
        L = fx.factor().to_dense_L()
    except Exception as e:
        print("trial", trial, "EXC", str(e)[-60:]); continue
    bad = np.argwhere(np.abs(L - Lref) > 1e-9)
    print("trial", trial, "nbad", len(bad))
    if len(bad):
        blocks = {}
        for r, c in bad[:200000]:
            blocks.setdefault((r // 128, c // 128), []).append((r % 128, c % 128))
        for (bi, bj), v in sorted(blocks.items())[:6]:
            rows = sorted(set(a for a, _ in v)); cols = sorted(set(b for _, b in v))
            print("  block", bi, bj, "count", len(v), "rows", rows[:6], "..", rows[-3:], "cols", cols[:6], "..", cols[-3:])
