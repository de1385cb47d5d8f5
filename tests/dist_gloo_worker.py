"""world_size-2 gloo worker (CPU): the host-side logic of the N>1 path -- NCCL-id rendezvous
plumbing, block-cyclic ownership partition, posterior row-chunk sharding."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import ctypes as C

import torch
import torch.distributed as dist


def main():
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group(backend="gloo", rank=rank, world_size=world)
    from stheno_jl_b200 import lib as sblib
    lib = sblib.load()
    ok = True
    # 1. id rendezvous: rank 0's 128-byte blob reaches every rank unchanged
    blob = [bytes(range(128)) if rank == 0 else None]
    dist.broadcast_object_list(blob, src=0)
    ok &= blob[0] == bytes(range(128))
    # 2. every trailing tile is updated by exactly one rank at every step
    for nblk in (1, 2, 7, 64, 512):
        for k in range(0, nblk, max(1, nblk // 9)):
            mine = torch.tensor([lib.sb_owned_trailing_tiles(nblk, k, rank, world)], dtype=torch.int64)
            dist.all_reduce(mine)
            t = nblk - k - 1
            ok &= mine.item() == t * (t + 1) // 2
            # my columns really are the ones owner() assigns to me
            cols = [J for J in range(k + 1, nblk) if lib.sb_owner_of_block(J, world) == rank]
            ok &= sum(nblk - J for J in cols) == lib.sb_owned_trailing_tiles(nblk, k, rank, world)
    # 3. row chunks partition [0, ns) contiguously in rank order
    for ns in (0, 1, 5, 4096, 4097):
        lo, hi = C.c_int64(), C.c_int64()
        assert lib.sb_row_chunk(ns, rank, world, C.byref(lo), C.byref(hi)) == 0
        g = [None] * world
        dist.all_gather_object(g, (lo.value, hi.value))
        ok &= g[0][0] == 0 and g[-1][1] == ns and all(g[i][1] == g[i + 1][0] for i in range(world - 1))
    t = torch.tensor([0 if ok else 1])
    dist.all_reduce(t)
    if rank == 0:
        print("GLOO_OK" if t.item() == 0 else "GLOO_FAIL")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
