"""Golden vectors (tests/golden/golden.json, 50-digit mpmath evaluation of the defining
formulas): the oracle (CPU suite) and the CUDA path (-m gpu) are both pinned to them."""
import json
import os

import numpy as np
import pytest

from models import f3_model

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = json.load(open(os.path.join(HERE, "golden", "golden.json")))
RTOL = 1e-10


def _run(m, case):
    if case["name"].startswith("se"):
        f = m.gppp(lambda GP: dict(f=GP(m.SEKernel())))
    else:
        f = f3_model(m)
    obs = m.BlockData(*[m.GPPPInput(p, np.array(x)) for p, x in case["blocks"]])
    tst = m.BlockData(*[m.GPPPInput(p, np.array(x)) for p, x in case["test_blocks"]])
    fx = f(obs, case["noise"])
    y = np.array(case["y"])
    lp = m.logpdf(fx, y)
    mean, var = m.mean_and_var(m.posterior(fx, y), tst)
    return lp, mean, var


def _check(lp, mean, var, case):
    np.testing.assert_allclose(lp, float(case["logpdf"]), rtol=RTOL)
    np.testing.assert_allclose(mean, [float(v) for v in case["mean"]], rtol=RTOL, atol=1e-12)
    np.testing.assert_allclose(var, [float(v) for v in case["var"]], rtol=1e-9, atol=1e-12)


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_matches_golden(orc, case):
    _check(*_run(orc, case), case)


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_cuda_path_matches_golden(sb, case):
    _check(*_run(sb, case), case)
