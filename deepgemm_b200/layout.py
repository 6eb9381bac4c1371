"""Scale-factor layout API (reference: csrc/apis/layout.hpp:14-122, csrc/jit_kernels/impls/smxx_layout.hpp:120-353).

The GEMM kernel consumes one wire format only -- the reference's SM100 format: UE8M0 exponent bytes, four consecutive
K granules packed into one int32, MN-major with the MN extent padded to 16 bytes:
``shape [.., mn, ceil(sf_k/4)]``, ``strides (.., 1, align(mn, 4))``.
"""
from typing import Optional, Sequence, Tuple, Union

import weakref

import torch

from ._lib import check, lib
from .runtime import get_mk_alignment_for_contiguous_layout, get_tma_aligned_size


def _require(cond: bool, what: str) -> None:
    if not cond:
        raise RuntimeError(f'Assertion error (deepgemm_b200/layout.py): {what}')


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def _empty_mn_major(num_batches: int, mn: int, cols: int, dtype, device) -> torch.Tensor:
    aligned_mn = get_tma_aligned_size(mn, 4)
    return torch.empty_strided((num_batches, mn, cols), (cols * aligned_mn, 1, aligned_mn), dtype=dtype, device=device)


def get_mn_major_tma_aligned_tensor(sf: torch.Tensor) -> torch.Tensor:
    """FP32 SFs -> MN-major, TMA-aligned FP32 (smxx_layout.hpp:120-153). Exported utility; the SM100 GEMM path does
    not consume FP32 scale factors (csrc/apis/gemm.hpp:118-122)."""
    _require(sf.dim() in (2, 3) and sf.dtype == torch.float32, 'sf must be a 2-D/3-D float tensor')
    batched = sf.unsqueeze(0) if sf.dim() == 2 else sf
    b, mn, sf_k = batched.shape
    aligned_mn = get_tma_aligned_size(mn, 4)
    if (batched.stride(0) == aligned_mn * sf_k or sf.dim() == 2) and batched.stride(1) == 1 and batched.stride(2) == aligned_mn:
        return sf
    out = _empty_mn_major(b, mn, sf_k, torch.float32, sf.device)
    check(lib().dgb200_transpose_sf_fp32(batched.data_ptr(), out.data_ptr(), mn, sf_k, b, batched.stride(0),
                                         batched.stride(1), batched.stride(2), _stream()))
    return out.squeeze(0) if sf.dim() == 2 else out


def _pack(sf: torch.Tensor, mn: int, gran_mn: int, psum_layout: Optional[torch.Tensor]) -> torch.Tensor:
    """Shared body: FP32 [.., ceil(mn/gran_mn), sf_k] (any strides) -> packed [.., mn, ceil(sf_k/4)]."""
    batched = sf.unsqueeze(0) if sf.dim() == 2 else sf
    b, rows, sf_k = batched.shape
    _require(rows == _ceil_div(mn, gran_mn), 'sf.size(-2) == ceil_div(mn, gran_mn)')
    out = _empty_mn_major(b, mn, _ceil_div(sf_k, 4), torch.int32, sf.device)
    if psum_layout is not None:
        _require(b == 1 and batched.is_contiguous(), 'psum layout needs one contiguous SF batch')
        _require(psum_layout.dtype == torch.int32 and psum_layout.is_contiguous() and psum_layout.numel() > 0,
                 'psum_layout must be a non-empty contiguous int tensor')
        check(lib().dgb200_pack_sf_ue8m0(batched.data_ptr(), out.data_ptr(), mn, sf_k, b, gran_mn, batched.stride(0),
                                         batched.stride(1), batched.stride(2), psum_layout.data_ptr(),
                                         psum_layout.numel(), get_mk_alignment_for_contiguous_layout(), _stream()))
    else:
        check(lib().dgb200_pack_sf_ue8m0(batched.data_ptr(), out.data_ptr(), mn, sf_k, b, gran_mn, batched.stride(0),
                                         batched.stride(1), batched.stride(2), None, 0, 1, _stream()))
    return out.squeeze(0) if sf.dim() == 2 else out


def get_mn_major_tma_aligned_packed_ue8m0_tensor(sf: torch.Tensor, psum_layout: Optional[torch.Tensor] = None) -> torch.Tensor:
    """FP32 power-of-two SFs [.., mn, sf_k] -> packed UE8M0 int32 [.., mn, ceil(sf_k/4)], strides (.., 1, align(mn,4))
    (smxx_layout.hpp:180-253; bit-exact contract pinned by the reference's tests/test_layout.py:20-42)."""
    _require(sf.dim() in (2, 3) and sf.dtype == torch.float32, 'sf must be a 2-D/3-D float tensor')
    return _pack(sf, sf.size(-2), 1, psum_layout)


def get_k_grouped_mn_major_tma_aligned_packed_ue8m0_tensor(sf: torch.Tensor, grouped_layout: torch.Tensor,
                                                           ks_cpu: Optional[Sequence[int]], gran_k: int, k_alignment: int,
                                                           use_psum_layout: bool = False) -> torch.Tensor:
    """K-grouped FP32 SFs [sum_g ceil(k_g/gran_k), mn] -> packed int32 [sum_g ceil(k_g/(4 gran_k)), mn]; every group is
    padded to a multiple of 4 granules on its own (smxx_layout.hpp:255-316). `grouped_layout` (device) holds per-group K,
    or end offsets with `use_psum_layout`; `ks_cpu` is only used to size the output (None / [] -> upper bound)."""
    _require(gran_k in (32, 128), 'gran_k == 32 or gran_k == 128')
    _require(k_alignment % 32 == 0, 'k_alignment % 32 == 0')
    _require(sf.dim() == 2 and sf.dtype == torch.float32 and sf.is_contiguous(), 'sf is a contiguous 2-D float tensor')
    sf_k, mn = sf.shape
    num_groups = grouped_layout.numel()
    _require(num_groups <= 128 and mn % 4 == 0, 'num_groups <= 128 and mn % 4 == 0')
    _require(grouped_layout.is_contiguous() and grouped_layout.dtype == torch.int32, 'grouped_layout is contiguous int32')
    has_ks = ks_cpu is not None and len(ks_cpu) > 0
    if has_ks:
        _require(len(ks_cpu) == num_groups, 'len(ks_cpu) == num_groups')
        packed_rows = sum(_ceil_div(k, gran_k * 4) for k in ks_cpu)
        _require(use_psum_layout or sum(_ceil_div(k, gran_k) for k in ks_cpu) == sf_k, 'sum(ceil(k/gran_k)) == sf.size(0)')
    else:
        _require(use_psum_layout, 'ks_cpu may only be omitted with use_psum_layout')
        packed_rows = (sf_k + num_groups * 3) // 4
    out = torch.empty((packed_rows, mn), dtype=torch.int32, device=sf.device)
    if packed_rows:
        check(lib().dgb200_pack_sf_ue8m0_k_grouped(sf.data_ptr(), out.data_ptr(), mn, grouped_layout.data_ptr(), num_groups,
                                                   packed_rows, gran_k, k_alignment if use_psum_layout else 0, _stream()))
    return out


def check_k_grouped_packed_ue8m0_tensor(sf: torch.Tensor, grouped_layout: torch.Tensor, ks_cpu, gran_k: int,
                                        k_alignment: int, use_psum_layout: bool) -> torch.Tensor:
    """Validate a caller-packed k-grouped SF tensor (smxx_layout.hpp:319-352)."""
    _require(sf.dtype == torch.int32 and sf.dim() == 2 and sf.is_contiguous(), 'sf is a contiguous 2-D int tensor')
    _require(sf.size(1) % 4 == 0 and sf.size(0) > 0, 'mn % 4 == 0 and packed_sf_k > 0')
    if ks_cpu is not None and len(ks_cpu) > 0:
        _require(len(ks_cpu) == grouped_layout.numel(), 'len(ks_cpu) == num_groups')
        if not use_psum_layout:
            _require(sf.size(0) >= sum(_ceil_div(k, gran_k * 4) for k in ks_cpu), 'enough packed SF rows')
    else:
        _require(use_psum_layout, 'ks_cpu may only be omitted with use_psum_layout')
    return sf


def transform_k_grouped_sf_into_required_layout(sf, ks_cpu, grouped_layout, recipe, k_alignment, use_psum_layout):
    """csrc/apis/layout.hpp:92-122 (arch 10 branch)."""
    _require(sf.dim() == 2 and recipe[0] == 1 and recipe[1] == 1, 'k-grouped SFs are 2-D with a (1, 1, gran_k) recipe')
    if sf.dtype == torch.float32:
        return get_k_grouped_mn_major_tma_aligned_packed_ue8m0_tensor(sf, grouped_layout, ks_cpu, recipe[2], k_alignment,
                                                                      use_psum_layout)
    if sf.dtype == torch.int32:
        return check_k_grouped_packed_ue8m0_tensor(sf, grouped_layout, ks_cpu, recipe[2], k_alignment, use_psum_layout)
    raise RuntimeError('Unknown cases')


def check_sf_layout(sf: torch.Tensor, mn: int, k: int, gran_mn: int, gran_k: int, num_groups: Optional[int],
                    tma_stride_check: bool = False, type_check: Optional[torch.dtype] = None) -> torch.Tensor:
    """csrc/utils/layout.hpp:80-117. (Shape and strides are read once: this sits on the per-call path of every GEMM.)"""
    dtype = sf.dtype
    if type_check is not None:
        _require(dtype == type_check, f'sf.dtype == {type_check}')
    is_float = dtype == torch.float32
    _require(is_float or dtype == torch.int32, 'sf must be float or int')
    shape, stride = sf.shape, sf.stride()
    _require(len(shape) == (3 if num_groups is not None else 2), 'sf.dim() == num_groups.has_value() + 2')
    if num_groups is not None:
        _require(shape[-3] == num_groups, 'sf.size(-3) == num_groups')
    _require(shape[-2] == -(-mn // gran_mn), 'sf.size(-2) == ceil_div(mn, gran_mn)')
    _require(shape[-1] == -(-k // (gran_k * (1 if is_float else 4))), 'sf.size(-1) == ceil_div(k, gran_k * (1 or 4))')
    if tma_stride_check:
        if num_groups is not None:
            _require(stride[-3] == stride[-1] * shape[-1], 'sf.stride(-3) == sf.stride(-1) * sf.size(-1)')
        _require(stride[-2] == 1 or mn == 1, 'sf must be MN-major')
        _require(stride[-1] == -(-mn // 4) * 4, 'sf.stride(-1) == tma_aligned(mn)')       # 4-byte elements: align(mn, 16 / 4)
    return sf


Recipe = Union[Tuple[int, int, int], Tuple[int, int]]


def transform_sf_into_required_layout(sf: torch.Tensor, mn: int, k: int, recipe: Recipe,
                                      num_groups: Optional[int] = None, is_sfa: Optional[bool] = None,
                                      disable_ue8m0_cast: bool = False,
                                      psum_layout: Optional[torch.Tensor] = None) -> torch.Tensor:
    """csrc/apis/layout.hpp:14-61, SM100 branch only (this library is sm_100a-only)."""
    recipe = tuple(recipe)
    if len(recipe) == 3:
        _require(is_sfa is not None, 'is_sfa must be given with a 3-tuple recipe')
        gran_mn, gran_k = (recipe[0] if is_sfa else recipe[1]), recipe[2]
    else:
        _require(len(recipe) == 2 and is_sfa is None, 'invalid recipe')
        gran_mn, gran_k = recipe
    if sf.dtype == torch.int32 and gran_mn == 1 and gran_k in (32, 128):
        # pre-packed scale factors (the per-call path of an inference loop): one pass of checks, no kernel
        return check_sf_layout(sf, mn, k, gran_mn, gran_k, num_groups, tma_stride_check=True, type_check=torch.int32)
    check_sf_layout(sf, mn, k, gran_mn, gran_k, num_groups)
    if sf.dtype == torch.float32 and gran_k in (32, 128):
        # The SM100 kernel needs power-of-two scales (hardware block scaling); the reference asserts the same
        # (layout.hpp:49-50) and leaves `disable_ue8m0_cast=True` without an SM100 kernel (gemm.hpp:118-122).
        _require(not disable_ue8m0_cast, 'not disable_ue8m0_cast (FP32 scale factors are cast to UE8M0 on SM100)')
        return _pack(sf, mn, gran_mn, psum_layout)
    raise RuntimeError('Unknown SF transformation')


def get_default_recipe(sfa_dtype: torch.dtype, sfb_dtype: torch.dtype) -> Tuple[int, int, int]:
    """csrc/utils/layout.hpp:64-77 (arch 10 branch)."""
    _require(sfb_dtype in (torch.float32, torch.int32), 'sfb must be float or int')
    return (1, 128, 128) if sfb_dtype == torch.float32 else (1, 1, 128)


_pair_memo = {}   # (id(sfa), id(sfb)) -> the last validation of that pair of PRE-PACKED scale-factor tensors (weak references +
                  # the metadata that was checked); a model has one pair per GEMM call site


def transform_sf_pair_into_required_layout(sfa, sfb, m, n, k, recipe, recipe_a, recipe_b, num_groups_a, num_groups_b,
                                           disable_ue8m0_cast=False, psum_layout=None):
    """csrc/apis/layout.hpp:63-90. Returns (sfa, sfb, gran_k_a, gran_k_b).

    An inference loop calls the GEMMs with the same pre-packed scale-factor tensors over and over (weights always, activations
    under CUDA graphs / static buffers); re-running the chain of layout checks costs ~4 us of a ~10 us kernel. Validated pairs
    are remembered by identity, and a repeat call only confirms that shapes and strides are still the ones that were checked
    (an in-place transpose or resize would change them)."""
    key = (id(sfa), id(sfb))
    memo = _pair_memo.get(key)
    args = (m, n, k, recipe, recipe_a, recipe_b, num_groups_a, num_groups_b)
    if (memo is not None and psum_layout is None and memo[0]() is sfa and memo[1]() is sfb and memo[2] == args
            and sfa.shape == memo[3] and sfa.stride() == memo[4] and sfb.shape == memo[5] and sfb.stride() == memo[6]):
        return sfa, sfb, memo[7], memo[8]
    out = _transform_sf_pair(sfa, sfb, m, n, k, recipe, recipe_a, recipe_b, num_groups_a, num_groups_b, disable_ue8m0_cast, psum_layout)
    if out[0] is sfa and out[1] is sfb and psum_layout is None:        # both were pre-packed: nothing was computed, only checked
        if len(_pair_memo) >= 1024:
            _pair_memo.clear()
        # (weak references only: the memo must not keep a caller's tensors alive)
        _pair_memo[key] = (weakref.ref(sfa), weakref.ref(sfb), args, sfa.shape, sfa.stride(), sfb.shape, sfb.stride(), out[2], out[3])
    return out


def _transform_sf_pair(sfa, sfb, m, n, k, recipe, recipe_a, recipe_b, num_groups_a, num_groups_b, disable_ue8m0_cast, psum_layout):
    if recipe_a is None and recipe is None:
        recipe = get_default_recipe(sfa.dtype, sfb.dtype)
    _require((recipe_a is None) == (recipe_b is None), 'recipe_a and recipe_b come as a pair')
    _require((recipe_a is None) != (recipe is None), "either 'recipe' or the 'recipe_a' + 'recipe_b' pair")
    if recipe is not None:
        tsfa = transform_sf_into_required_layout(sfa, m, k, recipe, num_groups_a, True, disable_ue8m0_cast, psum_layout)
        tsfb = transform_sf_into_required_layout(sfb, n, k, recipe, num_groups_b, False, disable_ue8m0_cast)
        return tsfa, tsfb, recipe[2], recipe[2]
    tsfa = transform_sf_into_required_layout(sfa, m, k, recipe_a, num_groups_a, None, disable_ue8m0_cast, psum_layout)
    tsfb = transform_sf_into_required_layout(sfb, n, k, recipe_b, num_groups_b, None, disable_ue8m0_cast)
    return tsfa, tsfb, recipe_a[1], recipe_b[1]
