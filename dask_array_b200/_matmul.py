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
fp32 mode).  float64 and 32/64-bit integer operands -- and N-d ``tensordot`` / ``einsum`` in general -- go
through ``BlockContract``: the same accumulate-over-k structure, on the exact SIMT GEMM
(``b2_gemm_tn_simt``) when the dtype has no tensor-core path.
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


class BlockContract(ArrayExpr):
    """``tensordot(a, b, axes=(la, lb))`` (``linalg/_tensordot.py:45-136``): output dims = the free axes of
    ``a`` then the free axes of ``b``; the contracted axes have identical chunks on both operands."""

    _parameters = ["a", "b", "la", "lb"]

    @property
    def chunks(self):
        a, b, la, lb = (self.operand(k) for k in ("a", "b", "la", "lb"))
        return tuple(c for d, c in enumerate(a.chunks) if d not in la) + tuple(c for d, c in enumerate(b.chunks) if d not in lb)

    @property
    def dtype(self):
        a, b = self.operand("a"), self.operand("b")
        if a.dtype == b.dtype and a.dtype in (np.dtype(np.float32), _bf16()):
            return np.dtype(np.float32)
        return np.promote_types(a.dtype, b.dtype)

    def _tree_label(self):
        return f"BlockContract(axes={self.operand('la')},{self.operand('lb')})"


_TENSOR_DTYPES = None


def _tensor_core(dt) -> bool:
    return dt in (np.dtype(np.float32), _bf16())


def tensordot(a, b, axes=2):
    """``tensordot`` (``linalg/_tensordot.py:45-136``) for N-d operands and any number of contracted axes.
    fp32 / bf16 run on the tensor cores; fp64 and integer operands on the exact SIMT GEMM."""
    from ._blockwise import Elemwise
    from ._collection import Array, asarray
    from ._rechunk import Rechunk

    a, b = asarray(a), asarray(b)
    if isinstance(axes, (int, np.integer)):
        la, lb = tuple(range(a.ndim - axes, a.ndim)), tuple(range(axes))
    else:
        la, lb = axes
        la = (la,) if isinstance(la, (int, np.integer)) else tuple(la)
        lb = (lb,) if isinstance(lb, (int, np.integer)) else tuple(lb)
    la, lb = tuple(x % a.ndim for x in la), tuple(x % b.ndim for x in lb)
    if len(la) != len(lb) or any(a.shape[i] != b.shape[j] for i, j in zip(la, lb)):
        raise ValueError("shape-mismatch for sum")
    ae, be = a.expr, b.expr
    if a.ndim == 2 and b.ndim == 2 and len(la) == 1 and a.dtype == b.dtype and _tensor_core(a.dtype):
        # the 2-D tensor-core case is the blocked matmul: operands read K-major as they lie
        aa = a if la[0] == 1 else a.T
        bb = b if lb[0] == 0 else b.T
        return matmul(aa, bb)
    dt = np.promote_types(a.dtype, b.dtype)
    if not (a.dtype == b.dtype and _tensor_core(a.dtype)):
        if dt.kind == "b":
            dt = np.dtype(np.int64)
        if dt.kind == "c" or dt.itemsize < 4 and dt.kind in "iu":
            raise NotImplementedError(f"tensordot of dtype {dt} has no B200 kernel (float32/64, bfloat16, 32/64-bit integers)")
        if dt == np.dtype(np.float16):
            dt = np.dtype(np.float32)
        if ae.dtype != dt:
            ae = Elemwise("astype", (ae,), (("dtype", dt.name),))
        if be.dtype != dt:
            be = Elemwise("astype", (be,), (("dtype", dt.name),))
    # align the contracted chunks (the blockwise alignment of the reference, ``unify_chunks``)
    want = list(be.chunks)
    for i, j in zip(la, lb):
        want[j] = ae.chunks[i]
    if tuple(want) != tuple(be.chunks):
        be = Rechunk(be, tuple(want))
    return Array(BlockContract(ae, be, la, lb))


def einsum(*operands, dtype=None, optimize=False, split_every=None, **kwargs):
    """``einsum`` (``_einsum.py:181-271``): operands are contracted pairwise, left to right, through
    ``tensordot``; labels that appear in one operand only and not in the output are summed first; the result
    is transposed into the requested label order.  Labels repeated inside one operand (diagonals) and batch
    labels (shared by two operands AND kept) have no B200 kernel and are refused."""
    from ._collection import asarray

    if not operands or not isinstance(operands[0], str):
        raise NotImplementedError("einsum: the subscripts-string form is supported")
    subs, ops = operands[0].replace(" ", ""), [asarray(o) for o in operands[1:]]
    if "." in subs:
        raise NotImplementedError("einsum with an ellipsis")
    if "->" in subs:
        ins, out = subs.split("->")
    else:
        ins = subs
        flat = ins.replace(",", "")
        out = "".join(sorted(c for c in set(flat) if flat.count(c) == 1))
    terms = ins.split(",")
    if len(terms) != len(ops):
        raise ValueError("einsum: number of subscripts does not match the operands")
    for t, o in zip(terms, ops):
        if len(t) != o.ndim:
            raise ValueError(f"einsum: operand has {o.ndim} dims but subscripts {t!r}")
        if len(set(t)) != len(t):
            raise NotImplementedError("einsum: a label repeated inside one operand (diagonal)")
    cur, lab = ops[0], terms[0]
    for k in range(1, len(ops)):
        nxt, nl = ops[k], terms[k]
        later = set(out).union(*[set(t) for t in terms[k + 1:]]) if k + 1 < len(terms) else set(out)

        def presum(x, xl, other):
            drop = tuple(i for i, c in enumerate(xl) if c not in other and c not in later)
            if drop:
                x = x.sum(axis=drop, split_every=split_every)
                xl = "".join(c for i, c in enumerate(xl) if i not in drop)
            return x, xl
        cur, lab = presum(cur, lab, nl)
        nxt, nl = presum(nxt, nl, lab)
        shared = [c for c in lab if c in nl]
        if any(c in later for c in shared):
            raise NotImplementedError("einsum: batch labels (shared by two operands and kept) have no B200 kernel")
        cur = tensordot(cur, nxt, axes=([lab.index(c) for c in shared], [nl.index(c) for c in shared]))
        lab = "".join(c for c in lab if c not in shared) + "".join(c for c in nl if c not in shared)
    drop = tuple(i for i, c in enumerate(lab) if c not in out)
    if drop:
        cur = cur.sum(axis=drop, split_every=split_every)
        lab = "".join(c for i, c in enumerate(lab) if i not in drop)
    if sorted(lab) != sorted(out):
        raise ValueError(f"einsum: output labels {out!r} are not produced by the operands")
    if lab != out:
        cur = cur.transpose(tuple(lab.index(c) for c in out))
    want = np.dtype(dtype) if dtype is not None else None
    if want is not None and cur.dtype != want:
        cur = cur.astype(want)
    return cur


def matmul(a, b):
    """``Array.__matmul__`` (``_collection.py:856``, ``linalg/_tensordot.py:253-334``) for 2-D operands: bf16 and
    fp32 on the tensor cores (``BlockGEMM``), every other number type through ``tensordot``'s exact path."""
    from ._collection import Array
    from ._rechunk import Rechunk

    if a.ndim == 0 or b.ndim == 0:
        raise ValueError("`matmul` does not support scalars.")
    if a.ndim > 2 or b.ndim > 2:
        raise NotImplementedError("B200 matmul handles 1-D and 2-D operands (stacked / broadcast batch matmul has no kernel)")
    if a.ndim == 1 or b.ndim == 1:
        if a.shape[-1] != b.shape[0]:
            raise ValueError(f"matmul: shapes {a.shape} and {b.shape} are not aligned")
        return tensordot(a, b, axes=((a.ndim - 1,), (0,)))
    if a.shape[1] != b.shape[0]:
        raise ValueError(f"matmul: shapes {a.shape} and {b.shape} are not aligned")
    if not (a.dtype == b.dtype and _tensor_core(a.dtype)):
        return tensordot(a, b, axes=((1,), (0,)))
    be = b.expr
    if isinstance(be, Transpose) and tuple(be.operand("axes")) == (1, 0):
        bt = be.operand("array")                       # x @ y.T : y is already (N, K)
    else:
        bt = Transpose(be, (1, 0))                     # materialised by the transpose kernel
    ae = a.expr
    if ae.chunks[1] != bt.chunks[1]:                   # align the contraction chunks (Elemwise._lower analogue)
        bt = Rechunk(bt, (bt.chunks[0], ae.chunks[1]))
    return Array(BlockGEMM(ae, bt))
