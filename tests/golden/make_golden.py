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


def case(name, blocks, kfun, noise, y, test_blocks, meanf=None):
    """blocks: list of (process, x array); kfun(p, q, a, b) -> mp value."""
    pts = [(p, mp.mpf(float(v))) for p, x in blocks for v in x]
    tps = [(p, mp.mpf(float(v))) for p, x in test_blocks for v in x]
    n, ns = len(pts), len(tps)
    K = mp.matrix(n, n)
    for i, (p, a) in enumerate(pts):
        for j, (q, b) in enumerate(pts):
            K[i, j] = kfun(p, q, a, b)
        K[i, i] += mp.mpf(float(noise[i])) if np.ndim(noise) else mp.mpf(noise)
    L = mp.cholesky(K)
    mf = meanf if meanf is not None else (lambda p, a: mp.mpf(0))
    yv = mp.matrix([mp.mpf(float(v)) - mf(p, a) for v, (p, a) in zip(y, pts)])
    logdet = 2 * sum(mp.log(L[i, i]) for i in range(n))
    v = mp.lu_solve(L, yv)  # L v = y
    quad = sum(v[i] ** 2 for i in range(n))
    logpdf = -(n * mp.log(2 * mp.pi) + logdet + quad) / 2
    alpha = mp.lu_solve(K, yv)
    mean, var = [], []
    for (p, a) in tps:
        kvec = mp.matrix([kfun(p, q, a, b) for (q, b) in pts])
        mean.append(mf(p, a) + sum(kvec[i] * alpha[i] for i in range(n)))
        w = mp.lu_solve(L, kvec)
        var.append(kfun(p, p, a, a) - sum(w[i] ** 2 for i in range(n)))
    return {
        "name": name, "noise": [float(v) for v in noise] if np.ndim(noise) else noise,
        "blocks": [[p, [float(v) for v in x]] for p, x in blocks],
        "test_blocks": [[p, [float(v) for v in x]] for p, x in test_blocks],
        "y": [float(v) for v in y],
        "logpdf": mp.nstr(logpdf, 25), "logdet": mp.nstr(logdet, 25),
        "mean": [mp.nstr(m, 25) for m in mean], "var": [mp.nstr(m, 25) for m in var],
    }


def k_m12(a, b):
    return mp.e ** (-abs(a - b))


def k_m32(a, b):
    s = mp.sqrt(3) * abs(a - b)
    return (1 + s) * mp.e ** (-s)


def k_white(a, b):
    return mp.mpf(1) if a == b else mp.mpf(0)


# ---- tests/models.py: rich_model, restated term by term (process = mean + sum_r c_r(x) a_r(g_r(x)))
def _m52_2d(u, v):
    d = mp.sqrt((u[0] - v[0]) ** 2 + (u[1] - v[1]) ** 2)
    s = mp.sqrt(5) * d
    return (1 + s + 5 * d * d / 3) * mp.e ** (-s)


RICH_ATOM_K = {
    "f1": lambda u, v: k_se(u, v),
    "f2": lambda u, v: mp.mpf("0.7") * k_m32(u, v) + mp.mpf("0.1") * k_white(u, v),
    "f3": lambda u, v: k_m12(u / 2, v / 2) + mp.mpf("0.3"),
    "f4": _m52_2d,
}
_one = lambda x: mp.mpf(1)
_id = lambda x: x
_per = lambda x: (mp.cos(2 * mp.pi * mp.mpf("0.25") * x), mp.sin(2 * mp.pi * mp.mpf("0.25") * x))
_sfun = lambda x: 1 + mp.mpf("0.1") * x * x
RICH = {  # name -> (mean(x), [(atom, coeff(x), inputmap(x))])
    "f1": (lambda x: mp.mpf(0), [("f1", _one, _id)]),
    "f2": (lambda x: mp.mpf("1.5"), [("f2", _one, _id)]),
    "f3": (lambda x: mp.cos(x), [("f3", _one, _id)]),
    "g1": (lambda x: mp.mpf(3), [("f1", _one, lambda x: mp.mpf("0.5") * x), ("f2", lambda x: mp.mpf(2), _id)]),
    "g2": (lambda x: -mp.cos(x), [("f1", _sfun, lambda x: x - mp.mpf("0.3")), ("f3", lambda x: mp.mpf(-1), _id)]),
    "g3": (lambda x: mp.sin(x) + mp.mpf("1.5"), [("f4", _one, _per), ("f2", _one, _id)]),
}
RICH["g4"] = (lambda x: RICH["g1"][0](x) - RICH["g2"][0](x),
              RICH["g1"][1] + [(a, (lambda c: (lambda x: -c(x)))(c), g) for a, c, g in RICH["g2"][1]])


def k_rich(p, q, a, b):
    v = mp.mpf(0)
    for (ap, cp, gp) in RICH[p][1]:
        for (aq, cq, gq) in RICH[q][1]:
            if ap == aq:
                v += cp(a) * cq(b) * RICH_ATOM_K[ap](gp(a), gq(b))
    return v


def vfe_case(name, x, z, kfun, noise, jitter, y):
    """AbstractGPs VFE: elbo, dtc for a single zero-mean process."""
    n, m = len(x), len(z)
    xm = [mp.mpf(float(v)) for v in x]
    zm = [mp.mpf(float(v)) for v in z]
    Kuu = mp.matrix(m, m)
    for i in range(m):
        for j in range(m):
            Kuu[i, j] = kfun(zm[i], zm[j])
        Kuu[i, i] += mp.mpf(jitter)
    Lu = mp.cholesky(Kuu)
    sig = mp.sqrt(mp.mpf(noise))
    Kuf = mp.matrix(m, n)
    for i in range(m):
        for j in range(n):
            Kuf[i, j] = kfun(zm[i], xm[j]) / sig
    A = mp.matrix(m, n)       # forward substitution: Lu A = Kuf / sigma  (A = U' \\ ...)
    for jj in range(n):
        for i in range(m):
            acc = Kuf[i, jj]
            for k in range(i):
                acc -= Lu[i, k] * A[k, jj]
            A[i, jj] = acc / Lu[i, i]
    D = A * A.T
    for i in range(m):
        D[i, i] += 1
    Ll = mp.cholesky(D)
    dt = mp.matrix([mp.mpf(float(v)) / sig for v in y])
    Ad = A * dt
    w = mp.matrix(m, 1)
    for i in range(m):
        acc = Ad[i]
        for k in range(i):
            acc -= Ll[i, k] * w[k]
        w[i] = acc / Ll[i, i]
    logdet_l = 2 * sum(mp.log(Ll[i, i]) for i in range(m))
    dd = sum(dt[i] ** 2 for i in range(n))
    dtc = -(n * mp.log(2 * mp.pi) + n * mp.log(mp.mpf(noise)) + logdet_l + dd - sum(w[i] ** 2 for i in range(m))) / 2
    tr = sum(kfun(xm[i], xm[i]) for i in range(n)) / mp.mpf(noise)
    fro = sum(A[i, j] ** 2 for i in range(m) for j in range(n))
    elbo = dtc - (tr - fro) / 2
    return {"name": name, "kind": "vfe", "x": [float(v) for v in x], "z": [float(v) for v in z], "noise": noise,
            "jitter": jitter, "y": [float(v) for v in y], "elbo": mp.nstr(elbo, 25), "dtc": mp.nstr(dtc, 25)}


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
    # Matern-1/2, single process, clustered inputs
    x = np.sort(rng.uniform(0, 4, 40))
    y = np.cos(x) + 0.2 * rng.standard_normal(40)
    out.append(case("matern12_n40", [("f", x)], lambda p, q, a, b: k_m12(a, b), 0.05, y, [("f", rng.uniform(0, 4, 5))]))
    # 0.7 Matern-3/2 + 0.1 White with coincident inputs (White is exact equality) and heteroscedastic noise
    x = rng.uniform(0, 6, 30)
    x = np.concatenate([x, x[:6]])
    noise = rng.uniform(0.05, 0.3, 36)
    y = rng.standard_normal(36)
    out.append(case("m32_white_hetero", [("f", x)],
                    lambda p, q, a, b: mp.mpf("0.7") * k_m32(a, b) + mp.mpf("0.1") * k_white(a, b), noise, y,
                    [("f", np.concatenate([rng.uniform(0, 6, 4), x[:2]]))]))
    # tests/models.py rich_model: every lowering rule (sum, function/const scaling, negation, stretch,
    # shift, periodic, kernel sums/scales, White + Constant kernels, constant / function means)
    rn = ["g1", "g4", "g3"]
    xr = [rng.uniform(-3, 3, k) for k in (12, 10, 11)]
    yr = rng.standard_normal(33) + 1.0
    xtr = [rng.uniform(-3, 3, 3) for _ in range(3)]
    out.append(case("rich_model", list(zip(rn, xr)), k_rich, 0.2, yr, list(zip(["g2", "f1", "g3"], xtr)),
                    meanf=lambda p, a: RICH[p][0](a)))
    # VFE: elbo / dtc with pseudo-points on a grid, and Z == X (elbo == exact logpdf, README.md:75-78)
    x = np.sort(rng.uniform(0, 10, 40))
    y = np.sin(x) + 0.3 * rng.standard_normal(40)
    out.append(vfe_case("vfe_se_m12", x, np.linspace(0, 10, 12), k_se, 0.1, 1e-9, y))
    out.append(vfe_case("vfe_se_z_eq_x", x, x, k_se, 0.1, 1e-9, y))
    out.append(case("se_n40_for_vfe", [("f", x)], lambda p, q, a, b: k_se(a, b), 0.1, y, [("f", x[:3])]))
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden.json")
    with open(path, "w") as fh:
        json.dump(out, fh, indent=1)
    print("wrote", path)


if __name__ == "__main__":
    main()
