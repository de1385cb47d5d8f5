"""Pins the oracle with every relation the reference's own tests hold for this path
(SURVEY.md section 4 / 8c): the reference has no golden numbers, only these identities."""
import numpy as np
import pytest

from models import f3_model, mixing_model, rich_model, toy_model


def test_atomic_boundary(orc):
    """test/gp/atomic_gp.jl:6-41."""
    rng = np.random.default_rng(0)
    x, xp = rng.standard_normal(7), rng.standard_normal(5)
    gpc = orc.GPC()
    k = orc.SEKernel()
    f = orc.atomic(orc.GP(np.sin, k), gpc)
    g = orc.atomic(orc.GP(orc.SEKernel()), gpc)
    assert np.array_equal(orc.mean(f, x), np.sin(x))                        # :14
    assert np.array_equal(orc.cov(f, x), orc.kernelmatrix(k, x))            # :15
    assert np.array_equal(orc.cov(f, x, xp, g=g), np.zeros((7, 5)))         # :33 independent atomics
    assert np.array_equal(orc.var(f, x), np.ones(7))                        # :34
    assert np.allclose(orc.cov(f, x, xp), orc.cov(f, xp, x).T)              # :36
    assert f.n == 1 and g.n == 2 and gpc.n == 2


def test_gppp_external_consistency(orc):
    """test/gaussian_process_probabilistic_programme.jl:27-43."""
    rng = np.random.default_rng(1)
    f = toy_model(orc)
    f1, f3 = f.fs["f1"], f.fs["f3"]
    x0, x1 = orc.GPPPInput("f1", rng.standard_normal(4)), orc.GPPPInput("f3", rng.standard_normal(3))
    assert np.array_equal(orc.mean(f1, x0.x), orc.mean(f, x0))
    assert np.array_equal(orc.cov(f3, x1.x), orc.cov(f, x1))
    assert np.array_equal(orc.cov(f1, x0.x, x1.x, g=f3), orc.cov(f, x0, x1))
    assert np.array_equal(orc.var(f3, x1.x, x0.x[:3], g=f1), orc.var(f, x1, orc.GPPPInput("f1", x0.x[:3])))
    z = rng.standard_normal(3)
    y = orc.rand(f(x1, 1e-3), z)
    a = orc.cov(orc.posterior(f3(x1.x, 1e-3), y), x1.x)
    b = orc.cov(orc.posterior(f(x1, 1e-3), y), x1)
    assert np.array_equal(a, b)                                             # :41-42


def _interface(orc, f, x0, x1, atol=1e-9):
    """AbstractGPs.TestUtils.test_internal_abstractgps_interface."""
    K = orc.cov(f, x0)
    n0 = len(x0) if not isinstance(x0, list) else len(x0)
    assert K.shape == (n0, n0)
    assert np.allclose(K, K.T, atol=atol)
    assert np.linalg.eigvalsh(K + 1e-9 * np.eye(n0)).min() > -1e-9
    assert np.allclose(orc.var(f, x0), np.diag(K), atol=atol)
    assert np.allclose(orc.cov(f, x0, x0), K, atol=atol)
    assert np.allclose(orc.cov(f, x0, x1), orc.cov(f, x1, x0).T, atol=atol)
    m, C = orc.mean_and_cov(f, x0)
    assert np.allclose(m, orc.mean(f, x0)) and np.allclose(C, K)


def test_gppp_internal_consistency_all_input_permutations(orc):
    """The nine (x0, x1) input-type permutations of gppp.jl test :47-86."""
    rng = np.random.default_rng(2)
    f = toy_model(orc)
    G, B = orc.GPPPInput, orc.BlockData
    r = rng.standard_normal
    cases = [
        (G("f1", r(4)), G("f3", r(3))),
        (G("f1", r(4)), B([G("f2", r(3)), G("f3", r(2))])),
        (B([G("f2", r(3)), G("f3", r(2))]), G("f1", r(4))),
        (B([G("f2", r(3)), G("f3", r(2))]), B([G("f1", r(6))])),
        (list(G("f1", r(4))), list(G("f3", r(3)))),
        (G("f1", r(4)), list(G("f3", r(3)))),
        (list(B([G("f2", r(3)), G("f3", r(2))])), list(G("f1", r(4)))),
        (list(B([G("f2", r(3)), G("f3", r(2))])), G("f1", r(4))),
        (B([list(G("f2", r(3))), G("f3", r(2))]), G("f1", r(4))),
    ]
    for x0, x1 in cases:
        _interface(orc, f, x0, x1)


def test_nested_gppp(orc):
    """gppp.jl test :107-120."""
    rng = np.random.default_rng(3)
    f = toy_model(orc)
    gpc = orc.GPC()
    f1o = orc.atomic(f, gpc)
    fo = orc.GPPP(dict(f1=f1o, f2=5 * f1o), gpc)
    x0 = orc.GPPPInput("f1", orc.GPPPInput("f1", rng.standard_normal(5)))
    x1 = orc.GPPPInput("f2", orc.GPPPInput("f2", rng.standard_normal(4)))
    _interface(orc, fo, x0, x1)


def test_cross_block_matrix(orc):
    """test/affine_transformations/cross.jl:55-76: block matrix == manual vcat/hcat."""
    rng = np.random.default_rng(4)
    f = f3_model(orc)
    xs = [rng.standard_normal(n) for n in (4, 3, 5)]
    names = ["f1", "f2", "f3"]
    K = orc.cov(f, orc.BlockData(*[orc.GPPPInput(n, x) for n, x in zip(names, xs)]))
    rows = []
    for ni, xi in zip(names, xs):
        rows.append(np.hstack([orc.cov(f.fs[ni], xi, xj, g=f.fs[nj]) for nj, xj in zip(names, xs)]))
    assert np.array_equal(K, np.vstack(rows))
    # B.1 worked trace: K13 = k1, K12 = 0, K33 = k1 + k2
    k1, k2 = orc.SEKernel(), orc.Matern52Kernel()
    assert np.array_equal(K[:4, 4:7], np.zeros((4, 3)))
    assert np.array_equal(K[:4, 7:], k1.matrix(xs[0], xs[2]))
    assert np.allclose(K[7:, 7:], k1.matrix(xs[2], xs[2]) + k2.matrix(xs[2], xs[2]), rtol=1e-15)


def test_addition_product_compose_rules(orc):
    """addition.jl:5-51, product.jl:16-44, compose.jl:11-21,53-54 of the reference tests."""
    rng = np.random.default_rng(5)
    x, xp = rng.standard_normal(6), rng.standard_normal(4)
    gpc = orc.GPC()
    f1, f2 = orc.atomic(orc.GP(np.sin, orc.SEKernel()), gpc), orc.atomic(orc.GP(np.cos, orc.SEKernel()), gpc)
    f3 = f1 + f2
    f4 = f1 + f3
    f5 = f3 + f4
    k = orc.SEKernel().matrix
    assert np.allclose(orc.cov(f3, x), 2 * k(x, x))
    assert np.allclose(orc.cov(f4, x), 5 * k(x, x))      # (2 f1 + f2): 4 k + k
    assert np.allclose(orc.cov(f5, x), 13 * k(x, x))     # (3 f1 + 2 f2): 9 k + 4 k
    assert np.allclose(orc.cov(f5, x, xp, g=f1), 3 * k(x, xp))
    g = 5.0 + f1                                          # known-function addition: exact ==
    assert np.array_equal(orc.cov(g, x), orc.cov(f1, x))
    assert np.array_equal(orc.mean(g, x), 5.0 + np.sin(x))
    h = 2.5 * (2.0 * f1)                                  # chained scalings
    assert np.allclose(orc.cov(h, x), 25.0 * k(x, x))
    assert np.allclose(orc.cov(h, x, xp, g=f1), 5.0 * k(x, xp))
    s = (lambda t: t * t) * f1
    assert np.allclose(orc.cov(s, x, xp), (x ** 2)[:, None] * k(x, xp) * (xp ** 2)[None, :])
    with pytest.raises(ValueError, match="Cannot multiply two GPs together"):
        f1 * f2
    c = orc.stretch(f1, 0.5)                              # cov(f o g)(x) == cov(f)(g.(x)) exactly
    assert np.array_equal(orc.cov(c, x), orc.cov(f1, 0.5 * x))
    assert np.all(np.diag(orc.cov(c, x, x)) == 1.0)       # k == 1.0 exactly at matched points
    assert np.array_equal(orc.cov(c, x, xp, g=f2), np.zeros((6, 4)))
    p = orc.periodic(f1, 0.5)
    assert np.allclose(orc.cov(p, x), orc.cov(p, x + 2.0), atol=1e-12)   # period 1/f
    assert np.array_equal(orc.cov(orc.shift(f1, 1.5), x), orc.cov(f1, x - 1.5))


def test_finite_gp_statistics_and_sparse(orc):
    """test/gp/util.jl:9-47 and test/gp/sparse_finite_gp.jl:11-42."""
    rng = np.random.default_rng(6)
    x = np.arange(0, 10.01, 0.1)
    xu = np.arange(0, 11.0)
    f = orc.atomic(orc.GP(orc.Matern32Kernel()), orc.GPC())
    fx = f(x, 1.0)
    fxu = orc.SparseFiniteGP(f(x, 1.0), f(xu, 1e-3))
    assert len(fxu) == len(x)
    with pytest.raises(RuntimeError, match="covariance matrix of a sparse GP"):
        orc.cov(fxu)
    z = rng.standard_normal(len(x))
    y = orc.rand(fx, z)
    p1 = orc.vfe_posterior(orc.VFE(fxu.finducing), fxu.fobs, y)
    p2 = orc.posterior(fxu, y)
    assert np.array_equal(orc.mean(p1, x), orc.mean(p2, x)) and np.array_equal(orc.var(p1, x), orc.var(p2, x))
    assert orc.elbo(fxu, y) == orc.logpdf(fxu, y) == orc.elbo(orc.VFE(fxu.finducing), fxu.fobs, y)
    for _ in range(10):                                                    # logpdf(fx, Y) .> elbo
        yy = orc.rand(fx, rng.standard_normal(len(x)))
        assert orc.logpdf(fx, yy) > orc.logpdf(fxu, yy)
    # README.md:75-78: pseudo-points == observations  =>  elbo == logpdf
    xs = rng.uniform(0, 5, 40)
    g = orc.atomic(orc.GP(orc.SEKernel()), orc.GPC())
    gx = g(xs, 0.1)
    yy = orc.rand(gx, rng.standard_normal(40))
    assert abs(orc.elbo(orc.VFE(g(xs, 1e-12)), gx, yy) - orc.logpdf(gx, yy)) < 1e-7
    # statistical check of rand (test/gp/util.jl:36-47, fewer samples)
    S = 20000
    Y = orc.rand(g(xs[:5], 0.0 + 1e-9), rng.standard_normal((5, S)))
    assert np.abs(Y.mean(axis=1)).max() < 5e-2
    assert np.abs(np.cov(Y) - orc.cov(g, xs[:5])).max() < 5e-2


def test_additive_gp_and_select(orc):
    rng = np.random.default_rng(7)
    X = rng.standard_normal((2, 9))
    gpc = orc.GPC()
    fa, fb = orc.atomic(orc.GP(orc.SEKernel()), gpc), orc.atomic(orc.GP(orc.Matern12Kernel()), gpc)
    f = orc.additive_gp([fa, fb])
    K = orc.cov(f, orc.ColVecs(X))
    assert np.allclose(K, orc.SEKernel().matrix(X[0]) + orc.Matern12Kernel().matrix(X[1]))


def test_fast_cpu_pipeline_equals_oracle(orc):
    """bench.py's timed CPU path computes the same numbers as the oracle entry points."""
    import bench
    x, y, xs = bench.make_inputs(300, 40)
    lp, m, v, _ = bench.cpu_pipeline(x, y, xs, 0.1)
    f = orc.gppp(lambda GP: dict(f=GP(orc.SEKernel())))
    fx = f(orc.GPPPInput("f", x), 0.1)
    assert abs(lp - orc.logpdf(fx, y)) < 1e-9 * abs(lp)
    mo, vo = orc.mean_and_var(orc.posterior(fx, y), orc.GPPPInput("f", xs))
    assert np.allclose(m, mo, rtol=1e-10, atol=1e-12) and np.allclose(v, vo, rtol=1e-9, atol=1e-12)
