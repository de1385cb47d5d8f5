"""Generate extended-precision golden vectors for the hot path (run: python tests/golden/make_golden.py).

The reference (Julia) cannot run here and its tests hold no golden numbers, so these fixtures pin
BOTH the oracle and the CUDA path to an independent 50-digit evaluation of the defining formulas
(AbstractGPs `logpdf`, `posterior` mean/var; KernelFunctions SE / Matern52 kernels with exact
(x-y)^2 distances) on the exact fp64 inputs stored alongside.
"""
import json
import os

import mpmath as mp
import numpy as np

mp.mp.dps = 50


def k_se(a, b):
    return mp.e ** (-(a - b) ** 2 / 2)


def k_m52(a, b):
    d = abs(a - b)
    s = mp.sqrt(5) * d
    return (1 + s + 5 * d * d / 3) * mp.e ** (-s)


def gram(ks, xs, ys):
    """ks[i][j] = kernel between process i and j (or None)."""
    return ks


def case(name, blocks, kfun, noise, y, test_blocks):
    """blocks: list of (process, x array); kfun(p, q, a, b) -> mp value."""
    pts = [(p, mp.mpf(float(v))) for p, x in blocks for v in x]
    tps = [(p, mp.mpf(float(v))) for p, x in test_blocks for v in x]
    n, ns = len(pts), len(tps)
    K = mp.matrix(n, n)
    for i, (p, a) in enumerate(pts):
        for j, (q, b) in enumerate(pts):
            K[i, j] = kfun(p, q, a, b)
        K[i, i] += mp.mpf(noise)
    L = mp.cholesky(K)
    yv = mp.matrix([mp.mpf(float(v)) for v in y])
    logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
    v = mp.lu_solve(L, yv)  # L v = y
    quad = sum(v[i] ** 2 for i in range(n))
    logpdf = -(n * mp.log(2 * mp.pi) + logdet + quad) / 2
    alpha = mp.lu_solve(K, yv)
    mean, var = [], []
    for (p, a) in tps:
        kvec = mp.matrix([kfun(p, q, a, b) for (q, b) in pts])
        mean.append(sum(kvec[i] * alpha[i] for i in range(n)))
        w = mp.lu_solve(L, kvec)
        var.append(kfun(p, p, a, a) - sum(w[i] ** 2 for i in range(n)))
    return {
        "name": name, "noise": noise,
        "blocks": [[p, [float(v) for v in x]] for p, x in blocks],
        "test_blocks": [[p, [float(v) for v in x]] for p, x in test_blocks],
        "y": [float(v) for v in y],
        "logpdf": mp.nstr(logpdf, 25), "logdet": mp.nstr(logdet, 25),
        "mean": [mp.nstr(m, 25) for m in mean], "var": [mp.nstr(m, 25) for m in var],
    }


def main():
    rng = np.random.default_rng(20260923)
    out = []
    # config-1 shaped: single SE process, x ~ U(0, 8), sigma2 = 0.1
    x = np.sort(rng.uniform(0, 8, 48))
    y = np.sin(x) + 0.3 * rng.standard_normal(48)
    xs = rng.uniform(0, 8, 7)
    out.append(case("se_n48", [("f", x)], lambda p, q, a, b: k_se(a, b), 0.1, y, [("f", xs)]))
    # GPPP f3 = f1 + f2 (SE + Matern52), observe all three, predict all three
    xb = [rng.uniform(0, 5, n) for n in (14, 11, 17)]
    yb = rng.standard_normal(42)
    xt = [rng.uniform(0, 5, 3) for _ in range(3)]

    def k3(p, q, a, b):
        has1 = lambda s: s in ("f1", "f3")
        has2 = lambda s: s in ("f2", "f3")
        v = mp.mpf(0)
        if has1(p) and has1(q):
            v += k_se(a, b)
        if has2(p) and has2(q):
            v += k_m52(a, b)
        return v

    names = ["f1", "f2", "f3"]
    out.append(case("gppp_f3", list(zip(names, xb)), k3, 0.1, yb, list(zip(names, xt))))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
