"""TEST-ONLY interpreter of a lowered `sb_covspec` (the structure handed to the C ABI), written
with the oracle's kernel functions.  It lets the CPU suite check the host-side plan lowering
(stheno.jl_b200/gp.py) against the oracle's recursive restatement of the reference routing
without a GPU.  It is not importable from the product package."""
import ctypes as C

import numpy as np

from oracle import stheno_oracle as o

_K = {0: o.SEKernel(), 1: o.Matern12Kernel(), 2: o.Matern32Kernel(), 3: o.Matern52Kernel(), 4: o.WhiteKernel()}


def _array(spec, i):
    a = spec.arrays[i]
    n, d = a.n, max(a.dim, 1)
    buf = (C.c_double * (n * d)).from_address(a.data)
    arr = np.frombuffer(buf, dtype=np.float64).reshape(n, d).copy()
    return arr if a.dim else arr[:, 0]


def _inp(z):
    return z[:, 0] if z.shape[1] == 1 else o.ColVecs(z.T)


def eval_dense(spec):
    K = np.zeros((spec.nrows, spec.ncols))
    for b in range(spec.nblocks):
        B = spec.blocks[b]
        blk = np.zeros((B.nrows, B.ncols))
        for t in range(B.term0, B.term0 + B.nterms):
            T = spec.terms[t]
            zl, zr = _array(spec, T.zl), _array(spec, T.zr)
            if T.kernel == 5:
                k = np.full((B.nrows, B.ncols), T.param)
            else:
                k = _K[T.kernel].matrix(_inp(zl), _inp(zr))
            sl = _array(spec, T.sl)[:, None] if T.sl >= 0 else 1.0
            sr = _array(spec, T.sr)[None, :] if T.sr >= 0 else 1.0
            blk += T.coeff * sl * sr * k
        K[B.row0:B.row0 + B.nrows, B.col0:B.col0 + B.ncols] = blk
    if spec.symmetric:
        K = np.tril(K) + np.tril(K, -1).T
    return K


def eval_diag(spec):
    v = np.zeros(spec.nrows)
    for b in range(spec.nblocks):
        B = spec.blocks[b]
        acc = np.zeros(B.nrows)
        for t in range(B.term0, B.term0 + B.nterms):
            T = spec.terms[t]
            zl, zr = _array(spec, T.zl), _array(spec, T.zr)
            if T.kernel == 5:
                k = np.full(B.nrows, T.param)
            else:
                k = _K[T.kernel].diag(_inp(zl), _inp(zr))
            sl = _array(spec, T.sl) if T.sl >= 0 else 1.0
            sr = _array(spec, T.sr) if T.sr >= 0 else 1.0
            acc += T.coeff * sl * sr * k
        v[B.row0:B.row0 + B.nrows] = acc
    return v
