"""Blocked matmul (``linalg/_tensordot.py:194-334``) -- lowered to the tcgen05 block GEMM."""
from __future__ import annotations


def matmul(a, b):
    raise NotImplementedError("blocked matmul on tcgen05 is not wired up yet (b2_gemm_tn)")
