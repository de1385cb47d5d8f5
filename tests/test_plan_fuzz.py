"""Fuzz the host-side plan lowering: random GP programmes (sums, scalar / function scalings,
negations, shifts, stretches, known-function additions over a few atomic leaves with assorted
kernels) must give the same block covariance / variance / mean as the oracle's recursive
restatement of the reference routing, for random BlockData layouts (including empty blocks)."""
import numpy as np
import pytest

from plan_eval import eval_dense, eval_diag

FUNCS = [np.sin, np.cos, lambda t: 1.0 + 0.1 * t * t, lambda t: 0.5 - 0.2 * t]


def _random_programme(m, rng, depth=5):
    """Build the SAME random programme in module `m` (product or oracle) from the rng stream."""
    def build(GP):
        kernels = [m.SEKernel(), m.Matern32Kernel(), 0.7 * m.Matern52Kernel() + 0.2 * m.WhiteKernel(),
                   m.with_lengthscale(m.Matern12Kernel(), 1.7) + m.ConstantKernel(0.25)]
        means = [None, 0.3, np.sin, None]
        leaves = [GP(k) if mu is None else GP(mu, k) for k, mu in zip(kernels, means)]
        procs = list(leaves)
        for _ in range(depth):
            op = rng.integers(0, 7)
            a = procs[rng.integers(0, len(procs))]
            b = procs[rng.integers(0, len(procs))]
            if op == 0:
                new = a + b
            elif op == 1:
                new = float(rng.uniform(-2, 2)) * a
            elif op == 2:
                new = FUNCS[rng.integers(0, len(FUNCS))] * a
            elif op == 3:
                new = a - b
            elif op == 4:
                new = m.shift(a, float(rng.uniform(-1, 1)))
            elif op == 5:
                new = m.stretch(a, float(rng.uniform(0.3, 2.0)))
            else:
                new = a + FUNCS[rng.integers(0, len(FUNCS))]
            procs.append(new)
        return {f"p{i}": p for i, p in enumerate(procs)}
    return m.gppp(build)


@pytest.mark.parametrize("seed", range(12))
def test_random_programmes(sb, orc, seed):
    from stheno_jl_b200.gp import Lowered, spec_dense, spec_diag, spec_symmetric
    fs = _random_programme(sb, np.random.default_rng(seed))
    fo = _random_programme(orc, np.random.default_rng(seed))
    # ids: every node (named or not) consumed one id from the shared counter, like the reference
    n_nodes = fs.gpc.n
    assert n_nodes == fo.gpc.n >= 9
    rng = np.random.default_rng(1000 + seed)
    names = list(fs.fs.keys())
    pick = [names[i] for i in rng.integers(0, len(names), 4)]
    sizes = [int(v) for v in rng.integers(0, 7, 4)]          # empty blocks allowed
    xs = [rng.uniform(-2, 2, n) for n in sizes]
    bs = sb.BlockData(*[sb.GPPPInput(p, x) for p, x in zip(pick, xs)])
    bo = orc.BlockData(*[orc.GPPPInput(p, x) for p, x in zip(pick, xs)])
    if len(bs) == 0:
        return
    lx = Lowered(fs, bs)
    Ko = orc.cov(fo, bo)
    assert np.allclose(eval_dense(spec_symmetric(lx)), Ko, rtol=1e-12, atol=1e-13)
    assert np.allclose(eval_diag(spec_diag(lx)), orc.var(fo, bo), rtol=1e-12, atol=1e-13)
    assert np.allclose(lx.mean(), orc.mean(fo, bo), rtol=1e-13, atol=1e-14)
    q = names[rng.integers(0, len(names))]
    xq = rng.uniform(-2, 2, 5)
    ly = Lowered(fs, sb.GPPPInput(q, xq))
    assert np.allclose(eval_dense(spec_dense(lx, ly)), orc.cov(fo, bo, orc.GPPPInput(q, xq)), rtol=1e-12, atol=1e-13)
    # unlike the reference (cross.jl:39 allocates a DerivedGP per `cov` call) indexing the
    # product's programme never mutates its GPC
    assert fs.gpc.n == n_nodes and fo.gpc.n > n_nodes
