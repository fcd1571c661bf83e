"""Blocked matmul (``linalg/_tensordot.py:253-334``).

The reference emits one ``np.matmul`` task per (i, k, j) block triple that keeps the contracted
axis as size 1 (``_matmul`` :194-213) and then sums the (M, 1, N) partials over k with
``reduce(np.add, ...)`` through a split_every tree (``_sum_wo_cat`` :216-249) -- 32 GiB of
partials at BASELINE config 5.  Here both steps are ONE node, ``BlockGEMM``: every output block
(i, j) is a single tcgen05 launch whose TMEM accumulator runs over all k blocks
(``b2_gemm_tn_pairs``).  Operands are taken K-major ("TN"): ``x @ y.T`` -- the config -- reads the
blocks of ``y`` as they are; a plain ``x @ w`` first materialises ``w.T`` with the transpose kernel.

Precision: bf16 operands -> exact products, fp32 accumulation (result dtype float32).  fp32
operands are split into three bf16 planes (hi + mid + lo, ~2^-24) and the six leading products
are accumulated, giving fp32-class accuracy (rtol 1e-5 against sgemm; tensor cores have no IEEE
fp32 mode).  Other dtypes raise NotImplementedError.
"""
from __future__ import annotations

import numpy as np

from ._blockwise import Transpose
from ._expr import ArrayExpr


def _bf16():
    import ml_dtypes

    return np.dtype(ml_dtypes.bfloat16)


class BlockGEMM(ArrayExpr):
    """``a @ bt.T`` with ``a`` (M, K) and ``bt`` (N, K), contraction chunks aligned."""

    _parameters = ["a", "bt"]

    @property
    def chunks(self):
        return (self.operand("a").chunks[0], self.operand("bt").chunks[0])

    @property
    def dtype(self):
        return np.dtype(np.float32)

    def _tree_label(self):
        return "BlockGEMM(tcgen05, k-accumulate)"


def matmul(a, b):
    """``Array.__matmul__`` (``_collection.py:856``) for 2-D operands."""
    from ._collection import Array
    from ._rechunk import Rechunk

    if a.ndim != 2 or b.ndim != 2:
        raise NotImplementedError("B200 matmul handles 2-D operands (the BASELINE contraction)")
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"matmul: shapes {a.shape} and {b.shape} are not aligned")
    ok = (np.dtype(np.float32), _bf16())
    if a.dtype not in ok or b.dtype != a.dtype:
        raise NotImplementedError(f"B200 matmul supports float32 @ float32 and bfloat16 @ bfloat16, got {a.dtype} @ {b.dtype}")
    be = b.expr
    if isinstance(be, Transpose) and tuple(be.operand("axes")) == (1, 0):
        bt = be.operand("array")                       # x @ y.T : y is already (N, K)
    else:
        bt = Transpose(be, (1, 0))                     # materialised by the transpose kernel
    ae = a.expr
    if ae.chunks[1] != bt.chunks[1]:                   # align the contraction chunks (Elemwise._lower analogue)
        bt = Rechunk(bt, (bt.chunks[0], ae.chunks[1]))
    return Array(BlockGEMM(ae, bt))
