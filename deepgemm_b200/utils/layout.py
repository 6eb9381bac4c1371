"""Layout re-exports (reference: deep_gemm/utils/layout.py:1-21)."""
from ..layout import (  # noqa: F401
    get_k_grouped_mn_major_tma_aligned_packed_ue8m0_tensor,
    get_mn_major_tma_aligned_packed_ue8m0_tensor,
    get_mn_major_tma_aligned_tensor,
    transform_sf_into_required_layout,
)
from ..runtime import (  # noqa: F401
    get_mk_alignment_for_contiguous_layout,
    get_theoretical_mk_alignment_for_contiguous_layout,
    get_tma_aligned_size,
    set_mk_alignment_for_contiguous_layout,
)
