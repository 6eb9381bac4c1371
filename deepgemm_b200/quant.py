"""CUDA activation quantiser: the step in front of the GEMM in inference.

`per_token_cast_to_fp8_packed(x)` produces, in ONE kernel, what the reference's callers build with
`per_token_cast_to_fp8(x, use_ue8m0=True, gran_k, use_packed_ue8m0=True)` (deep_gemm/utils/math.py:26-38; ~10 eager
torch kernels) followed by the GEMM's own scale-factor transform (csrc/apis/layout.hpp:48-58; transpose + pack
kernels): FP8 E4M3 rows and packed UE8M0 scale factors already in the MN-major, TMA-aligned wire format, so the
GEMM call that follows launches nothing but the GEMM. Bit-identical to the reference's Python (tests/test_quant_gpu.py).
There is no CPU fallback: the torch-only `deepgemm_b200.utils.per_token_cast_to_fp8` is the oracle-side restatement.
"""
from typing import Tuple

import torch

from ._lib import check, lib
from .runtime import get_tma_aligned_size


def per_token_cast_to_fp8_packed(x: torch.Tensor, gran_k: int = 128) -> Tuple[torch.Tensor, torch.Tensor]:
    """x [M, K] BF16 (row pitch a multiple of 8 elements for the vector path; any pitch works) ->
    (x_fp8 [M, K] float8_e4m3fn, sf int32 [M, ceil(K / (4 gran_k))] with strides (1, align(M, 4)))."""
    if not x.is_cuda:
        raise RuntimeError('per_token_cast_to_fp8_packed needs a CUDA tensor (use deepgemm_b200.utils.per_token_cast_to_fp8 on CPU)')
    if x.dim() != 2 or x.dtype != torch.bfloat16 or x.stride(1) != 1:
        raise RuntimeError('Assertion error (deepgemm_b200/quant.py): x is a 2-D BF16 tensor with contiguous rows')
    if gran_k not in (32, 128):
        raise RuntimeError('Assertion error (deepgemm_b200/quant.py): gran_k == 32 or gran_k == 128')
    m, k = x.shape
    q = torch.empty((m, k), dtype=torch.float8_e4m3fn, device=x.device)
    words = -(-k // (4 * gran_k))
    aligned_m = get_tma_aligned_size(m, 4)
    sf = torch.empty_strided((m, words), (1, aligned_m), dtype=torch.int32, device=x.device)
    if m == 0 or k == 0:
        return q, sf
    check(lib().dgb200_per_token_cast_to_fp8(x.data_ptr(), x.stride(0), q.data_ptr(), q.stride(0), sf.data_ptr(), aligned_m,
                                             m, k, gran_k, torch.cuda.current_stream().cuda_stream))
    return q, sf
