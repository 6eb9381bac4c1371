"""Parity metrics (same definitions as the reference's deep_gemm/testing/numeric.py:5-22)."""
import torch


def calc_diff(x: torch.Tensor, y: torch.Tensor) -> float:
    """1 - 2<x,y> / (|x|^2 + |y|^2) in FP64: 0 for identical tensors, the reference's test tolerance is < 1e-3."""
    x, y = x.double(), y.double()
    denom = (x * x + y * y).sum()
    if denom == 0:
        return 0.0
    return float(1 - 2 * (x * y).sum() / denom)


def count_bytes(*tensors) -> int:
    total = 0
    for t in tensors:
        if isinstance(t, (tuple, list)):
            total += count_bytes(*t)
        elif t is not None:
            total += t.numel() * t.element_size()
    return total
