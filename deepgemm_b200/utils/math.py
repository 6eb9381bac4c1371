"""Quantisation helpers callers use to build FP8 operands + scale factors.

Same contracts as the reference's ``deep_gemm/utils/math.py`` (:5-61): amax/448 scale per 1x`gran_k` (tokens),
`gran_k`x`gran_k` (weights) or `gran_k`x1 (wgrad) block, optionally rounded UP to a power of two (UE8M0), values
cast to E4M3 round-to-nearest. Written against torch only so they run on CPU (oracle side) and CUDA alike.
"""
from typing import Tuple

import torch

FP8_E4M3_MAX = 448.0


def ceil_div(x: int, y: int) -> int:
    return -(-x // y)


def align(x: int, y: int) -> int:
    return ceil_div(x, y) * y


def ceil_to_ue8m0(x: torch.Tensor) -> torch.Tensor:
    """Smallest power of two >= |x| with the exponent clamped to [1, 254] (math.py:13-16)."""
    raw = x.abs().float().contiguous().view(torch.int32)
    exponent = (raw >> 23) & 0xFF
    exponent = exponent + ((raw & 0x7FFFFF) != 0).to(torch.int32)
    return (exponent.clamp_(1, 254) << 23).view(torch.float32)


def pack_ue8m0_to_int(x: torch.Tensor) -> torch.Tensor:
    """FP32 powers of two [.., 4j] -> int32 [.., j], byte i of a word = exponent of element 4j+i (math.py:19-23)."""
    assert x.dtype == torch.float32 and x.size(-1) % 4 == 0
    raw = x.contiguous().view(torch.int32)
    assert bool((raw >= 0).all()) and bool(((raw & 0x7FFFFF) == 0).all()), 'scale factors must be powers of two'
    return (raw >> 23).to(torch.uint8).view(torch.int32)


def unpack_ue8m0_from_int(packed: torch.Tensor) -> torch.Tensor:
    return (packed.contiguous().view(torch.uint8).to(torch.int32) << 23).view(torch.float32)


def _scale_from_amax(amax: torch.Tensor, use_ue8m0: bool) -> torch.Tensor:
    sf = amax.clamp(min=1e-4) / FP8_E4M3_MAX
    return ceil_to_ue8m0(sf) if use_ue8m0 else sf


def per_token_cast_to_fp8(x: torch.Tensor, use_ue8m0: bool, gran_k: int = 128,
                          use_packed_ue8m0: bool = False) -> Tuple[torch.Tensor, torch.Tensor]:
    """1 x gran_k scaling along the last dim of a 2-D tensor (math.py:26-38)."""
    assert x.dim() == 2
    m, k = x.shape
    k_pad = align(k, gran_k)
    xp = x.new_zeros((m, k_pad))
    xp[:, :k] = x
    blocks = xp.view(m, k_pad // gran_k, gran_k)
    sf = _scale_from_amax(blocks.abs().float().amax(dim=2), use_ue8m0)
    q = (blocks * (1.0 / sf.unsqueeze(2))).to(torch.float8_e4m3fn).view(m, k_pad)[:, :k].contiguous()
    return q, (pack_ue8m0_to_int(sf) if use_packed_ue8m0 else sf)


def per_channel_cast_to_fp8(x: torch.Tensor, use_ue8m0: bool, gran_k: int = 128) -> Tuple[torch.Tensor, torch.Tensor]:
    """gran_k x 1 scaling along the first dim (wgrad operands, math.py:41-48)."""
    assert x.dim() == 2 and x.size(0) % gran_k == 0
    k, n = x.shape
    blocks = x.view(k // gran_k, gran_k, n)
    sf = _scale_from_amax(blocks.abs().float().amax(dim=1), use_ue8m0)
    q = (blocks * (1.0 / sf.unsqueeze(1))).to(torch.float8_e4m3fn).view(k, n)
    return q, sf


def per_block_cast_to_fp8(x: torch.Tensor, use_ue8m0: bool, gran_k: int = 128) -> Tuple[torch.Tensor, torch.Tensor]:
    """gran_k x gran_k scaling (weights, math.py:51-61)."""
    assert x.dim() == 2
    m, n = x.shape
    mp, np_ = align(m, gran_k), align(n, gran_k)
    xp = x.new_zeros((mp, np_))
    xp[:m, :n] = x
    blocks = xp.view(mp // gran_k, gran_k, np_ // gran_k, gran_k)
    sf = _scale_from_amax(blocks.abs().float().amax(dim=(1, 3), keepdim=True), use_ue8m0)
    q = (blocks * (1.0 / sf)).to(torch.float8_e4m3fn).view(mp, np_)[:m, :n].contiguous()
    return q, sf.view(mp // gran_k, np_ // gran_k)


def per_custom_dims_cast_to_fp8(x: torch.Tensor, dims: Tuple, use_ue8m0: bool) -> Tuple[torch.Tensor, torch.Tensor]:
    reduce_dims = tuple(i for i in range(x.dim()) if i not in set(dims))
    sf = _scale_from_amax(x.abs().float().amax(dim=reduce_dims, keepdim=True), use_ue8m0)
    return (x * (1.0 / sf)).to(torch.float8_e4m3fn), sf.squeeze()
