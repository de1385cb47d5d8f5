"""2-rank VFE/elbo check (run under torch.distributed.run): chunk-sharded accumulation + all-reduce
must reproduce the oracle's elbo / dtc / approximate posterior."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

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
    ids = [sblib.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(ids, src=0)
    ctx = sblib.Context(local, rank, world, ids[0])
    sblib.set_default_context(ctx)
    rng = np.random.default_rng(3)
    n, m = 40000, 200   # 3 chunks of 16384 rows -> uneven split over 2 ranks
    x = rng.uniform(0, 30, n)
    z = np.linspace(0, 30, m)
    xs = rng.uniform(0, 30, 77)
    y = np.sin(x) + 0.3 * rng.standard_normal(n)
    fs, fo = sb.gppp(lambda GP: dict(f=GP(sb.SEKernel()))), orc.gppp(lambda GP: dict(f=GP(orc.SEKernel())))
    fxs, fzs = fs(sb.GPPPInput("f", x), 0.1), fs(sb.GPPPInput("f", z), 1e-6)
    fxo, fzo = fo(orc.GPPPInput("f", x), 0.1), fo(orc.GPPPInput("f", z), 1e-6)
    e, eo = sb.elbo(sb.VFE(fzs), fxs, y), orc.elbo(orc.VFE(fzo), fxo, y)
    ps, po = sb.posterior(sb.SparseFiniteGP(fxs, fzs), y), orc.posterior(orc.SparseFiniteGP(fxo, fzo), y)
    mm = sb.mean(ps, sb.GPPPInput("f", xs))
    ok = abs(e - eo) <= 1e-9 * abs(eo) and np.allclose(mm, orc.mean(po, orc.GPPPInput("f", xs)), rtol=1e-7, atol=1e-8)
    t = torch.tensor([0 if ok else 1])
    dist.all_reduce(t)
    if rank == 0:
        print("VFE_DIST_OK" if t.item() == 0 else f"VFE_DIST_FAIL {e} {eo}")
    dist.barrier()
    sblib.set_default_context(None)
    ctx.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
