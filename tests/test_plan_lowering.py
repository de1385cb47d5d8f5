"""Host logic without a GPU: the product's flattened term plan (what crosses the C ABI) must
reproduce the oracle's *recursive* restatement of the reference routing
(src/gp/derived_gp.jl:31-59) for every model / input-type combination."""
import numpy as np
import pytest

from models import f3_model, mixing_model, rich_model, toy_model
from plan_eval import eval_dense, eval_diag


def _pair(sb, orc, builder):
    return builder(sb), builder(orc)


@pytest.mark.parametrize("builder,names", [
    (f3_model, ["f1", "f2", "f3"]), (toy_model, ["f3", "f1", "f2"]),
    (rich_model, ["g1", "g4", "g3", "f2", "g2"]), (mixing_model, ["g1", "g5", "g3", "f1"]),
])
def test_plan_matches_recursion(sb, orc, builder, names):
    from stheno_jl_b200.gp import Lowered, spec_dense, spec_diag, spec_symmetric
    rng = np.random.default_rng(11)
    fs, fo = _pair(sb, orc, builder)
    xs = [rng.uniform(-2, 2, 5 + 2 * i) for i in range(len(names))]
    bs = sb.BlockData(*[sb.GPPPInput(n, x) for n, x in zip(names, xs)])
    bo = orc.BlockData(*[orc.GPPPInput(n, x) for n, x in zip(names, xs)])
    lx = Lowered(fs, bs)
    Ko = orc.cov(fo, bo)
    assert np.allclose(eval_dense(spec_symmetric(lx)), Ko, rtol=1e-13, atol=1e-14)
    assert np.allclose(eval_dense(spec_dense(lx, lx)), Ko, rtol=1e-13, atol=1e-14)
    assert np.allclose(eval_diag(spec_diag(lx)), orc.var(fo, bo), rtol=1e-13, atol=1e-14)
    assert np.allclose(lx.mean(), orc.mean(fo, bo), rtol=1e-14, atol=1e-15)
    # cross-covariance against a single other process
    xo = rng.uniform(-2, 2, 6)
    ly = Lowered(fs, sb.GPPPInput(names[1], xo))
    assert np.allclose(eval_dense(spec_dense(lx, ly)), orc.cov(fo, bo, orc.GPPPInput(names[1], xo)),
                       rtol=1e-13, atol=1e-14)


def test_plan_b1_trace(sb):
    """f3 = f1 + f2 -> 6 lower blocks, terms {11:k1, 22:k2, 31:k1, 32:k2, 33:k1+k2},
    block (2,1) empty (the reference computes 8 kernel matrices + 8 dense zeros)."""
    from stheno_jl_b200.gp import Lowered, spec_symmetric
    f = f3_model(sb)
    x = np.linspace(0, 1, 4)
    spec = spec_symmetric(Lowered(f, sb.BlockData(*[sb.GPPPInput(n, x) for n in ("f1", "f2", "f3")])))
    assert spec.nblocks == 6
    nterms = [spec.blocks[i].nterms for i in range(6)]
    assert nterms == [1, 0, 1, 1, 1, 2]
    kinds = [[spec.terms[t].kernel for t in range(spec.blocks[i].term0, spec.blocks[i].term0 + spec.blocks[i].nterms)] for i in range(6)]
    assert kinds == [[0], [], [3], [0], [3], [0, 3]]
    assert spec.narrays == 1  # the shared input vector is uploaded once


def test_plan_b2_mixing_coefficients(sb):
    """cov(g_i, g_j) = a_i a_j k1 + b_i b_j k2 (2 terms per tile)."""
    from stheno_jl_b200.gp import Lowered, spec_dense
    f = mixing_model(sb)
    x = np.linspace(0, 1, 3)
    spec = spec_dense(Lowered(f, sb.GPPPInput("g1", x)), Lowered(f, sb.GPPPInput("g2", x)))
    assert spec.nterms == 2
    c = sorted(spec.terms[t].coeff for t in range(2))
    assert np.allclose(c, sorted([0.2 * 0.3, 0.8 * 0.7]))


def test_tuple_vector_regrouping(sb, orc):
    """gppp.jl:32-43: a vector of (symbol, feature) tuples is regrouped by first occurrence."""
    from stheno_jl_b200.gp import Lowered, spec_dense
    rng = np.random.default_rng(12)
    fs, fo = _pair(sb, orc, toy_model)
    pts = [("f3", 0.1), ("f1", 0.2), ("f3", 0.3), ("f2", 0.4), ("f1", 0.5)]
    ls = Lowered(fs, pts)
    assert [p is fs.fs[n] for p, n in zip(ls.procs, ["f3", "f1", "f2"])] == [True] * 3
    assert ls.lengths == [2, 2, 1]
    assert np.allclose(eval_dense(spec_dense(ls, ls)), orc.cov(fo, pts), rtol=1e-13, atol=1e-14)


def test_errors_match_reference(sb):
    gpc1, gpc2 = sb.GPC(), sb.GPC()
    a, b = sb.atomic(sb.GP(sb.SEKernel()), gpc1), sb.atomic(sb.GP(sb.SEKernel()), gpc2)
    with pytest.raises(AssertionError):
        a + b                                           # addition.jl:9  @assert fa.gpc === fb.gpc
    with pytest.raises(ValueError, match="Cannot multiply two GPs together"):
        a * a                                           # product.jl:13
    with pytest.raises(RuntimeError, match="covariance matrix of a sparse GP"):
        sb.cov(sb.SparseFiniteGP(a(np.zeros(3)), a(np.zeros(2))))   # sparse_finite_gp.jl:39-43
    assert a.n == 1 and gpc1.n == 1
    c = 2 * a + 1.0
    assert c.n == 3 and gpc1.n == 3                     # unnamed intermediates still get ids


def test_nested_gppp_lowering(sb, orc):
    """test/gaussian_process_probabilistic_programme.jl:107-120: a GPPP used as an atomic inside
    another GPPP; inputs GPPPInput(:f1, GPPPInput(:f1, x))."""
    from stheno_jl_b200.gp import Lowered, spec_dense, spec_diag
    rng = np.random.default_rng(13)

    def build(m):
        inner = toy_model(m)
        gpc = m.GPC()
        f1 = m.atomic(inner, gpc)
        f3 = m.atomic(inner, gpc)            # a second, independent wrapper of the same programme
        return m.GPPP(dict(f1=f1, f2=5 * f1, f3=f3), gpc)

    fs, fo = build(sb), build(orc)
    x0, x1 = rng.standard_normal(5), rng.standard_normal(4)
    for (a, ia), (b, ib) in [(("f1", "f1"), ("f2", "f2")), (("f2", "f3"), ("f1", "f1")), (("f1", "f3"), ("f3", "f3"))]:
        ls = Lowered(fs, sb.GPPPInput(a, sb.GPPPInput(ia, x0)))
        lt = Lowered(fs, sb.GPPPInput(b, sb.GPPPInput(ib, x1)))
        K = eval_dense(spec_dense(ls, lt))
        Ko = orc.cov(fo, orc.GPPPInput(a, orc.GPPPInput(ia, x0)), orc.GPPPInput(b, orc.GPPPInput(ib, x1)))
        assert np.allclose(K, Ko, rtol=1e-13, atol=1e-14)
        assert np.allclose(ls.mean(), orc.mean(fo, orc.GPPPInput(a, orc.GPPPInput(ia, x0))), rtol=1e-14)
        assert np.allclose(eval_diag(spec_diag(ls)), orc.var(fo, orc.GPPPInput(a, orc.GPPPInput(ia, x0))), rtol=1e-13)
    # different outer wrappers of the same inner programme are independent => no terms at all
    assert spec_dense(Lowered(fs, sb.GPPPInput("f1", sb.GPPPInput("f1", x0))),
                      Lowered(fs, sb.GPPPInput("f3", sb.GPPPInput("f1", x1)))).nterms == 0
