"""Multi-GPU worker (run under torch.distributed.run, one rank per GPU): the 1-D block-cyclic
Cholesky with NCCL panel broadcast + row-sharded posterior must reproduce the oracle."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np
import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    import stheno_jl_b200 as sb
    from stheno_jl_b200 import lib as sblib
    from oracle import stheno_oracle as orc
    from models import f3_model
    ids = [sblib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = sblib.Context(local, rank, world, ids[0])
    sblib.set_default_context(ctx)

    rng = np.random.default_rng(5)
    fails = []
    for n, ns in [(1000, 77), (3000, 333), (130, 5)]:
        x = rng.uniform(0, n / 32 + 1, n)
        xs = rng.uniform(0, n / 32 + 1, ns)
        y = np.sin(x) + 0.3 * rng.standard_normal(n)
        fs = sb.gppp(lambda GP: dict(f=GP(sb.SEKernel())))
        fo = orc.gppp(lambda GP: dict(f=GP(orc.SEKernel())))
        fxs, fxo = fs(sb.GPPPInput("f", x), 0.1), fo(orc.GPPPInput("f", x), 0.1)
        lp, lpo = sb.logpdf(fxs, y), orc.logpdf(fxo, y)
        ps, po = sb.posterior(fxs, y), orc.posterior(fxo, y)
        m, v = sb.mean_and_var(ps, sb.GPPPInput("f", xs))
        mo, vo = orc.mean_and_var(po, orc.GPPPInput("f", xs))
        L = fxs.factor().to_dense_L()
        import scipy.linalg as sla
        Lref = sla.cholesky(orc.cov(fxo), lower=True)
        ok = (abs(lp - lpo) <= 1e-10 * abs(lpo) and np.allclose(m, mo, rtol=1e-10, atol=1e-11)
              and np.allclose(v, vo, rtol=1e-10, atol=1e-11) and np.allclose(L, Lref, rtol=0, atol=1e-12))
        if not ok:
            fails.append((n, float(lp), float(lpo), float(np.abs(m - mo).max()), float(np.abs(L - Lref).max())))
    # GPPP with ragged blocks
    fs, fo = f3_model(sb), f3_model(orc)
    xs3 = [rng.uniform(0, 20, k) for k in (700, 513, 300)]
    bs = sb.BlockData(*[sb.GPPPInput(nm, x) for nm, x in zip(["f1", "f2", "f3"], xs3)])
    bo = orc.BlockData(*[orc.GPPPInput(nm, x) for nm, x in zip(["f1", "f2", "f3"], xs3)])
    y = orc.rand(fo(bo, 0.1), np.random.default_rng(1).standard_normal(1513))
    lp, lpo = sb.logpdf(fs(bs, 0.1), y), orc.logpdf(fo(bo, 0.1), y)
    if abs(lp - lpo) > 1e-10 * abs(lpo):
        fails.append(("gppp", float(lp), float(lpo)))
    t = torch.tensor([len(fails)], dtype=torch.int64)
    dist.all_reduce(t)
    if rank == 0:
        print("DIST_OK" if t.item() == 0 else f"DIST_FAIL {fails}")
    ctx.close()
    dist.destroy_process_group()
    sys.exit(0 if t.item() == 0 else 1)


if __name__ == "__main__":
    main()
