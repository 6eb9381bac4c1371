"""deepgemm_b200 -- B200-native FP8 blockwise-scaled GEMM behind the ``deep_gemm`` API.

Drop-in scope: fp8_gemm_{nt,nn,tn,tt}, m_grouped_fp8_gemm_{nt,nn}_contiguous,
m_grouped_fp8_gemm_nt_masked, k_grouped_fp8_gemm_tn_contiguous, fp8_gemm_nt_skip_head_mid, fp8_einsum, the SF layout
transforms and the runtime knobs; plus the CUDA activation quantiser `per_token_cast_to_fp8_packed`.
Importing this package does not touch CUDA (the reference guarantees the same, tests/test_lazy_init.py).
"""
from . import testing, utils  # noqa: F401
from .gemm import (  # noqa: F401
    bf16_gemm_nn, bf16_gemm_nt, bf16_gemm_tn, bf16_gemm_tt, einsum, k_grouped_bf16_gemm_tn_contiguous,
    m_grouped_bf16_gemm_nn_contiguous, m_grouped_bf16_gemm_nt_contiguous, m_grouped_bf16_gemm_nt_masked,
    fp8_bmm, fp8_einsum, fp8_gemm_nn, fp8_gemm_nt, fp8_gemm_nt_skip_head_mid, fp8_gemm_tn, fp8_gemm_tt,
    k_grouped_fp8_gemm_nt_contiguous, k_grouped_fp8_gemm_tn_contiguous,
    m_grouped_fp8_gemm_nn_contiguous, m_grouped_fp8_gemm_nt_contiguous, m_grouped_fp8_gemm_nt_masked,
)
from .layout import (  # noqa: F401
    get_k_grouped_mn_major_tma_aligned_packed_ue8m0_tensor,
    get_mn_major_tma_aligned_packed_ue8m0_tensor,
    get_mn_major_tma_aligned_tensor,
    transform_sf_into_required_layout,
)
from .quant import per_token_cast_to_fp8_packed  # noqa: F401
from .runtime import (  # noqa: F401
    get_mk_alignment_for_contiguous_layout, get_num_sms, get_pdl, get_split_k, get_tc_util,
    get_theoretical_mk_alignment_for_contiguous_layout, get_tma_aligned_size,
    set_block_size_multiple_of, set_ignore_compile_dims, set_mk_alignment_for_contiguous_layout,
    set_num_sms, set_pdl, set_split_k, set_tc_util,
)

# canonical names of the reference (csrc/apis/gemm.hpp:649-717): `fp8_fp4_*`, with `fp8_*` as aliases
fp8_fp4_gemm_nt, fp8_fp4_gemm_nn, fp8_fp4_gemm_tn, fp8_fp4_gemm_tt = fp8_gemm_nt, fp8_gemm_nn, fp8_gemm_tn, fp8_gemm_tt
m_grouped_fp8_fp4_gemm_nt_contiguous = m_grouped_fp8_gemm_nt_contiguous
m_grouped_fp8_fp4_gemm_nn_contiguous = m_grouped_fp8_gemm_nn_contiguous
m_grouped_fp8_fp4_gemm_nt_masked = m_grouped_fp8_gemm_nt_masked
fp8_m_grouped_gemm_nt_masked = m_grouped_fp8_gemm_nt_masked
bf16_m_grouped_gemm_nt_masked = m_grouped_bf16_gemm_nt_masked

__version__ = '0.1.0'
