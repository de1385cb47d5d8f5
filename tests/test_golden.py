"""Golden vectors (tests/golden/golden.json, 50-digit mpmath evaluation of the defining
formulas): the oracle (CPU suite) and the CUDA path (-m gpu) are both pinned to them."""
import json
import os

import numpy as np
import pytest

from models import f3_model, rich_model

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "golden.json")))
RTOL = 1e-10


def _model(m, case):
    name = case["name"]
    if name.startswith("se") or name.startswith("vfe_se"):
        return m.gppp(lambda GP: dict(f=GP(m.SEKernel())))
    if name.startswith("matern12"):
        return m.gppp(lambda GP: dict(f=GP(m.Matern12Kernel())))
    if name.startswith("m32_white"):
        return m.gppp(lambda GP: dict(f=GP(0.7 * m.Matern32Kernel() + 0.1 * m.WhiteKernel())))
    if name == "rich_model":
        return rich_model(m)
    return f3_model(m)


def _run(m, case):
    f = _model(m, case)
    if case.get("kind") == "vfe":
        x, z, y = np.array(case["x"]), np.array(case["z"]), np.array(case["y"])
        fx, fz = f(m.GPPPInput("f", x), case["noise"]), f(m.GPPPInput("f", z), case["jitter"])
        return m.elbo(m.VFE(fz), fx, y), m.dtc(m.VFE(fz), fx, y)
    obs = m.BlockData(*[m.GPPPInput(p, np.array(x)) for p, x in case["blocks"]])
    tst = m.BlockData(*[m.GPPPInput(p, np.array(x)) for p, x in case["test_blocks"]])
    noise = np.array(case["noise"]) if isinstance(case["noise"], list) else case["noise"]
    fx = f(obs, noise)
    y = np.array(case["y"])
    lp = m.logpdf(fx, y)
    mean, var = m.mean_and_var(m.posterior(fx, y), tst)
    return lp, mean, var


def _check(res, case):
    if case.get("kind") == "vfe":
        # K_uu + 1e-9 I has condition ~1e10: fp64 evaluations of the bound agree with the 50-digit
        # value to ~cond * eps
        np.testing.assert_allclose(res[0], float(case["elbo"]), rtol=2e-6)
        np.testing.assert_allclose(res[1], float(case["dtc"]), rtol=2e-6)
        return
    lp, mean, var = res
    np.testing.assert_allclose(lp, float(case["logpdf"]), rtol=RTOL)
    np.testing.assert_allclose(mean, [float(v) for v in case["mean"]], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(var, [float(v) for v in case["var"]], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(orc, case):
    _check(_run(orc, case), case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_path_matches_golden(sb, case):
    _check(_run(sb, case), case)
