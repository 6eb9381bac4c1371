"""Python face of the FP8 GEMM path -- same names, argument meaning and error behaviour as the reference's
``deep_gemm`` module (csrc/apis/gemm.hpp:73-400, signatures registered at :649-717).

Every function validates like the reference (RuntimeError on violation), converts scale factors to the packed
UE8M0 wire format when FP32 ones are given, and forwards raw device pointers to the C ABI (include/dgb200.h) on the
current torch CUDA stream. Nothing here computes: without the CUDA library the call fails.
"""
from typing import List, Optional, Tuple

import torch

from . import layout as _layout
from ._lib import check, lib
from .runtime import get_mk_alignment_for_contiguous_layout

_K_MAJOR, _MN_MAJOR = 0, 1
_BF16, _FP32 = 0, 1

TensorPair = Tuple[torch.Tensor, torch.Tensor]


def _require(cond: bool, what: str) -> None:
    if not cond:
        raise RuntimeError(f'Assertion error (deepgemm_b200/gemm.py): {what}')


def _stream() -> int:
    return torch._C._cuda_getCurrentRawStream(torch.cuda.current_device())


# Split-K scratch (include/dgb200.h, `workspace`): one zero-initialised buffer per (device, stream), so GEMMs that may
# run concurrently never share it. Only small problems use it (fewer output tiles than SM pairs).
_WORKSPACE_CAP = 64 << 20
_workspaces = {}


def _workspace(m: int, n: int, device: torch.device, stream: int):
    need = 16384 + 32 * m * n
    if need > _WORKSPACE_CAP:
        return None, 0
    key = (device.index, stream)
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < need:
        if torch.cuda.is_current_stream_capturing():
            return None, 0                      # never allocate inside a CUDA-graph capture
        ws = torch.zeros(max(need, 8 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws.data_ptr(), ws.numel()


def _major_check(t: torch.Tensor) -> int:
    """csrc/utils/layout.hpp:13-19; returns the major of the checked tensor."""
    shape, stride = t.shape, t.stride()
    _require(len(shape) in (2, 3), 'dim == 2 or dim == 3')
    if len(shape) == 3:
        _require(stride[0] == shape[-2] * shape[-1], 't.stride(0) == t.size(-2) * t.size(-1)')
    _require(stride[-2] == 1 or stride[-1] == 1, 't.stride(-2) == 1 or t.stride(-1) == 1')
    return _K_MAJOR if stride[-1] == 1 else _MN_MAJOR


def _major_ab(t: torch.Tensor) -> int:
    return _major_check(t)


def _check_cd(t: torch.Tensor) -> None:
    _require(_major_check(t) == _K_MAJOR, 'C/D must be row-major (stride(-1) == 1)')


def _check_fp8(t: torch.Tensor) -> None:
    if t.dtype != torch.float8_e4m3fn:
        if t.dtype == torch.int8:
            raise RuntimeError('FP4 (int8-packed) operands are outside this library\'s FP8xFP8 scope')
        raise RuntimeError(f'Assertion error (deepgemm_b200/gemm.py): operand dtype must be torch.float8_e4m3fn, got {t.dtype}')


def _d_dtype(d: torch.Tensor) -> int:
    _require(d.dtype in (torch.bfloat16, torch.float32), 'd.dtype is bfloat16 or float')
    return _BF16 if d.dtype == torch.bfloat16 else _FP32


def _early_return(m: int, n: int, k: int, d: torch.Tensor, c: Optional[torch.Tensor]) -> bool:
    """csrc/apis/gemm.hpp:19-46."""
    if m == 0 or n == 0:
        return True
    same = c is not None and c.data_ptr() == d.data_ptr()
    if same:
        _require(c.shape == d.shape and c.stride() == d.stride(), 'c and d alias with different layouts')
    _d_dtype(d)
    if c is not None:
        _check_cd(c)
        _require(d.dtype == c.dtype, 'd.dtype == c.dtype')
    if k == 0:
        if not same:
            d.copy_(c) if c is not None else d.zero_()
        return True
    if c is not None and not same:
        d.copy_(c)
    return False


# ------------------------------------------------------------------------------------------------ dense
def fp8_gemm_nt(a: TensorPair, b: TensorPair, d: torch.Tensor, c: Optional[torch.Tensor] = None,
                recipe: Optional[Tuple[int, int, int]] = None, recipe_a: Optional[Tuple[int, int]] = None,
                recipe_b: Optional[Tuple[int, int]] = None, compiled_dims: str = 'nk',
                disable_ue8m0_cast: bool = False) -> None:
    """D = (C +) A @ B.T with A [M,K], B [N,K] FP8 E4M3 and per-(1 x gran_k | 128 x 128) scale factors.
    Reference: fp8_fp4_gemm_nt, csrc/apis/gemm.hpp:73-124. `compiled_dims` is a JIT hint there; ignored here."""
    (a_t, sfa), (b_t, sfb) = a, b
    _check_fp8(a_t), _check_fp8(b_t)
    sha, shb, shd = a_t.shape, b_t.shape, d.shape
    _require(len(sha) == 2 and len(shb) == 2 and len(shd) == 2, 'a, b, d are 2-D')
    sta, stb, std = a_t.stride(), b_t.stride(), d.stride()
    _require(sta[0] == 1 or sta[1] == 1, 't.stride(-2) == 1 or t.stride(-1) == 1')
    _require(stb[0] == 1 or stb[1] == 1, 't.stride(-2) == 1 or t.stride(-1) == 1')
    _require(std[1] == 1, 'C/D must be row-major (stride(-1) == 1)')
    major_a = _K_MAJOR if sta[1] == 1 else _MN_MAJOR
    major_b = _K_MAJOR if stb[1] == 1 else _MN_MAJOR
    (m, k), (n, k_), (m_, n_) = sha, shb, shd
    _require(m == m_ and n == n_ and k == k_, 'm == m_ and n == n_ and k == k_')
    d_dtype = _d_dtype(d)
    if _early_return(m, n, k, d, c):
        return
    sfa_t, sfb_t, gran_k_a, gran_k_b = _layout.transform_sf_pair_into_required_layout(
        sfa, sfb, m, n, k, recipe, recipe_a, recipe_b, None, None, disable_ue8m0_cast)
    _require(sfa_t.dtype == torch.int32 and sfb_t.dtype == torch.int32, 'Unsupported architecture or scaling factor types')
    lda = sta[0] if major_a == _K_MAJOR else sta[1]
    ldb = stb[0] if major_b == _K_MAJOR else stb[1]
    stream = _stream()
    ws_ptr, ws_bytes = _workspace(m, n, d.device, stream)
    check(lib().dgb200_fp8_gemm_nt(a_t.data_ptr(), sfa_t.data_ptr(), b_t.data_ptr(), sfb_t.data_ptr(), d.data_ptr(),
                                   m, n, k, lda, ldb, std[0], major_a, major_b,
                                   sfa_t.stride(-1), sfb_t.stride(-1), gran_k_a, gran_k_b, d_dtype,
                                   int(c is not None), ws_ptr, ws_bytes, stream))


def fp8_gemm_nt_skip_head_mid(a: TensorPair, b: TensorPair, d: torch.Tensor, head_splits: Tuple[int, int, int],
                              recipe: Optional[Tuple[int, int, int]] = None, compiled_dims: str = 'nk',
                              disable_ue8m0_cast: bool = False) -> None:
    """D[:, head h] = [ left | (mid columns left untouched) | right ] of A @ B.T's head h: the GEMM's N is a multiple of
    left + right, D is n + n / (left + right) * mid wide and GEMM column j lands at j + (j + right) // (left + right) * mid.
    Used by DeepSeek-V3 MLA to write the no-PE / PE halves of every head around a gap filled elsewhere.
    Reference: csrc/apis/attention.hpp:19-74 (epilogue remap deep_gemm/include/deep_gemm/epilogue/transform.cuh:15-22)."""
    (a_t, sfa), (b_t, sfb) = a, b
    _check_fp8(a_t), _check_fp8(b_t)
    _require(_major_ab(a_t) == _K_MAJOR and _major_ab(b_t) == _K_MAJOR, 'major_a == K and major_b == K')
    _check_cd(d)
    _require(a_t.dim() == 2 and b_t.dim() == 2 and d.dim() == 2, 'a, b, d are 2-D')
    (m, k), (n, k_), (m_, n_) = a_t.shape, b_t.shape, d.shape
    _require(m == m_ and k == k_, 'm == m_ and k == k_')
    _require(n > 0 and k > 0, 'n > 0 and k > 0')
    _d_dtype(d)
    left, mid, right = (int(x) for x in head_splits)
    _require(left >= 0 and mid >= 0 and right >= 0 and left + right > 0, 'head splits are non-negative')
    _require(n % (left + right) == 0 and n_ == n + n // (left + right) * mid, 'n % (left + right) == 0 and n_ == n + n / (left + right) * mid')
    if m == 0:
        return
    sfa_t, sfb_t, gran_k_a, gran_k_b = _layout.transform_sf_pair_into_required_layout(
        sfa, sfb, m, n, k, recipe, None, None, None, None, disable_ue8m0_cast)
    _require(gran_k_a == 128 and gran_k_b == 128, 'gran_k_a == 128 and gran_k_b == 128')
    _require(sfa_t.dtype == torch.int32 and sfb_t.dtype == torch.int32, 'Unsupported architecture or scaling factor types')
    check(lib().dgb200_fp8_gemm_nt_skip_head_mid(
        a_t.data_ptr(), sfa_t.data_ptr(), b_t.data_ptr(), sfb_t.data_ptr(), d.data_ptr(), m, n, k, a_t.stride(0), b_t.stride(0),
        d.stride(0), left, mid, right, sfa_t.stride(-1), sfb_t.stride(-1), _d_dtype(d), _stream()))


def fp8_bmm(a: torch.Tensor, sfa: torch.Tensor, b: torch.Tensor, sfb: torch.Tensor, d: torch.Tensor,
            c: Optional[torch.Tensor] = None, recipe: Optional[Tuple[int, int, int]] = None,
            compiled_dims: str = 'nk') -> None:
    """D[i] = (C[i] +) A[i] @ B[i].T over a batch, A [B, M, K], B [B, N, K], D [B, M, N]; every tensor may be a permuted
    view (only the innermost-stride rules of a GEMM operand apply), which is what `fp8_einsum` feeds it.
    Reference: csrc/apis/einsum.hpp:137-175 (sm100_fp8_bmm, csrc/jit_kernels/impls/sm100_fp8_fp4_gemm_1d1d.hpp:393-467)."""
    _check_fp8(a), _check_fp8(b)
    _require(a.dim() == 3 and b.dim() == 3 and d.dim() == 3, 'a, b, d are 3-D')
    _require(a.stride(-1) == 1 or a.stride(-2) == 1, 'a.stride(-1) == 1 or a.stride(-2) == 1')
    _require(b.stride(-1) == 1 or b.stride(-2) == 1, 'b.stride(-1) == 1 or b.stride(-2) == 1')
    _require(d.stride(-1) == 1, 'd.stride(-1) == 1')
    major_a = _K_MAJOR if a.stride(-1) == 1 else _MN_MAJOR
    major_b = _K_MAJOR if b.stride(-1) == 1 else _MN_MAJOR
    (bs, m, k), (bs_, n, k_), (bs__, m_, n_) = a.shape, b.shape, d.shape
    _require(bs == bs_ == bs__, 'batch sizes agree')
    _require(m == m_ and n == n_ and k == k_, 'm == m_ and n == n_ and k == k_')
    _d_dtype(d)
    if bs == 0 or _early_return(m, n, k, d, c):
        return
    sfa_t, sfb_t, gran_k_a, gran_k_b = _layout.transform_sf_pair_into_required_layout(
        sfa, sfb, m, n, k, recipe, None, None, bs, bs, False)
    _require(sfa_t.dtype == torch.int32 and sfb_t.dtype == torch.int32, 'Unsupported architecture or scaling factor types')
    lda = a.stride(1) if major_a == _K_MAJOR else a.stride(2)
    ldb = b.stride(1) if major_b == _K_MAJOR else b.stride(2)
    check(lib().dgb200_fp8_bmm(a.data_ptr(), sfa_t.data_ptr(), b.data_ptr(), sfb_t.data_ptr(), d.data_ptr(), bs, m, n, k,
                               lda, ldb, d.stride(1), a.stride(0), b.stride(0), d.stride(0), major_a, major_b,
                               sfa_t.stride(-1), sfb_t.stride(-1), gran_k_a, gran_k_b, _d_dtype(d), int(c is not None),
                               _stream()))


def fp8_einsum(expr: str, a: TensorPair, b: TensorPair, d: torch.Tensor, c: Optional[torch.Tensor] = None,
               recipe: Tuple[int, int, int] = (1, 128, 128)) -> None:
    """The three hard-wired FP8 contractions of the reference (csrc/apis/einsum.hpp:177-214), each a batched GEMM over
    permuted views -- (batch, m, n, k) = (h, b, d, r), (h, b, r, d) and (h, d, r, b) -- with no copies."""
    (a_t, sfa), (b_t, sfb) = a, b
    if expr == 'bhr,hdr->bhd':
        fp8_bmm(a_t.permute(1, 0, 2), sfa.permute(1, 0, 2), b_t, sfb, d.permute(1, 0, 2),
                None if c is None else c.permute(1, 0, 2), recipe, 'nk')
    elif expr == 'bhd,hdr->bhr':
        fp8_bmm(a_t.permute(1, 0, 2), sfa.permute(1, 0, 2), b_t.permute(0, 2, 1), sfb.permute(0, 2, 1), d.permute(1, 0, 2),
                None if c is None else c.permute(1, 0, 2), recipe, 'nk')
    elif expr == 'bhd,bhr->hdr':
        fp8_bmm(a_t.permute(1, 2, 0), sfa.permute(1, 2, 0), b_t.permute(1, 2, 0), sfb.permute(1, 2, 0), d, c, recipe, 'mn')
    else:
        raise RuntimeError(f'Unsupported einsum expression: {expr}')


# ------------------------------------------------------------------------------------------------ BF16 operands
def _bf16_major(t: torch.Tensor, what: str):
    """(major, pitch of the strided extent in elements) of a BF16 operand's last two dims (get_major_type_ab, csrc/utils/layout.hpp)."""
    _require(t.dtype == torch.bfloat16, f'{what}.dtype == bfloat16')
    if t.stride(-1) == 1:
        major, contiguous, pitch = _K_MAJOR, t.size(-1), t.stride(-2)
    else:
        _require(t.stride(-2) == 1, f'{what} must be K-major or MN-major')
        major, contiguous, pitch = _MN_MAJOR, t.size(-2), t.stride(-1)
    _require(contiguous % 8 == 0 and pitch % 8 == 0 and t.data_ptr() % 16 == 0, f'{what}: 16-byte aligned rows')
    return major, pitch


def bf16_gemm_nt(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, c: Optional[torch.Tensor] = None,
                 compiled_dims: str = 'nk') -> None:
    """D = (C +) A @ B.T with BF16 A [M,K], B [N,K] (no scale factors; either may be a transposed view, i.e. MN-major),
    D BF16 or FP32. Reference: bf16_gemm_nt, csrc/apis/gemm.hpp:404-438 (kernel impls/sm100_bf16_gemm.cuh)."""
    _require(a.dim() == 2 and b.dim() == 2 and d.dim() == 2, 'a, b, d are 2-D')
    (major_a, lda), (major_b, ldb) = _bf16_major(a, 'a'), _bf16_major(b, 'b')
    _check_cd(d)
    (m, k), (n, k_), (m_, n_) = a.shape, b.shape, d.shape
    _require(m == m_ and n == n_ and k == k_, 'm == m_ and n == n_ and k == k_')
    d_dtype = _d_dtype(d)
    if _early_return(m, n, k, d, c):
        return
    check(lib().dgb200_bf16_gemm_nt(a.data_ptr(), b.data_ptr(), d.data_ptr(), m, n, k, lda, ldb, d.stride(0),
                                    major_a, major_b, d_dtype, int(c is not None), _stream()))


def bf16_gemm_nn(a, b, d, c=None, compiled_dims='nk'):
    """B given as [K,N] (gemm.hpp:440-446)."""
    bf16_gemm_nt(a, b.transpose(0, 1), d, c, compiled_dims)


def bf16_gemm_tn(a, b, d, c=None, compiled_dims='mn'):
    """A given as [K,M], B as [K,N] (gemm.hpp:448-454)."""
    bf16_gemm_nt(a.transpose(0, 1), b.transpose(0, 1), d, c, compiled_dims)


def bf16_gemm_tt(a, b, d, c=None, compiled_dims='mn'):
    """A given as [K,M] (gemm.hpp:456-462)."""
    bf16_gemm_nt(a.transpose(0, 1), b, d, c, compiled_dims)


def m_grouped_bf16_gemm_nt_contiguous(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, grouped_layout: torch.Tensor,
                                      compiled_dims: str = 'nk', use_psum_layout: bool = False,
                                      ensure_zero_padding: bool = True,
                                      expected_m_for_psum_layout: Optional[int] = None) -> None:
    """A [M_sum,K] BF16 rows grouped by expert, B [G,N,K] BF16 (or a transposed view of [G,K,N]), D [M_sum,N] BF16
    (gemm.hpp:464-517)."""
    _require(a.dim() == 2 and b.dim() == 3 and d.dim() == 2, 'a 2-D, b 3-D, d 2-D')
    (major_a, lda), (major_b, ldb) = _bf16_major(a, 'a'), _bf16_major(b, 'b')
    _require(major_a == _K_MAJOR, 'a is K-major')                                  # gemm.hpp:478
    _require(b.stride(0) == b.size(1) * b.size(2) and ldb == (b.size(2) if major_b == _K_MAJOR else b.size(1)), 'b is densely batched')
    _require(grouped_layout.is_contiguous() and grouped_layout.dtype == torch.int32, 'grouped_layout is contiguous int32')
    (m, k), (num_groups, n, k_), (m_, n_) = a.shape, b.shape, d.shape
    _require(m == m_ and n == n_ and k == k_, 'm == m_ and n == n_ and k == k_')
    _require(n > 0 and k > 0 and num_groups > 0, 'n > 0 and k > 0 and num_groups > 0')
    _require(d.dtype == torch.bfloat16, 'd.dtype == bfloat16')
    if use_psum_layout:
        _require(grouped_layout.dim() == 1 and grouped_layout.numel() == num_groups, 'grouped_layout is [num_groups]')
    else:
        _require(grouped_layout.dim() == 1 and grouped_layout.numel() == m, 'grouped_layout is [m]')
        _require(expected_m_for_psum_layout is None, 'expected_m_for_psum_layout needs use_psum_layout')
    _check_cd(d)
    if m == 0:
        return
    check(lib().dgb200_m_grouped_bf16_gemm_nt_contiguous(
        a.data_ptr(), b.data_ptr(), d.data_ptr(), grouped_layout.data_ptr(), num_groups, m, n, k, lda, ldb, d.stride(0), major_b,
        int(use_psum_layout), int(ensure_zero_padding), -1 if expected_m_for_psum_layout is None else int(expected_m_for_psum_layout),
        _stream()))


def m_grouped_bf16_gemm_nn_contiguous(a, b, d, grouped_layout, compiled_dims='nk', use_psum_layout=False, ensure_zero_padding=True):
    """B given as [G,K,N] (gemm.hpp:519-526)."""
    m_grouped_bf16_gemm_nt_contiguous(a, b.transpose(1, 2), d, grouped_layout, compiled_dims, use_psum_layout, ensure_zero_padding, None)


def m_grouped_bf16_gemm_nt_masked(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, masked_m: torch.Tensor, expected_m: int,
                                  compiled_dims: str = 'nk') -> None:
    """A [G,M_max,K], B [G,N,K] BF16, D [G,M_max,N] BF16; rows >= masked_m[g] are not written (gemm.hpp:528-564)."""
    _require(a.dim() == 3 and b.dim() == 3 and d.dim() == 3, 'a, b, d are 3-D')
    _require(a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, 'a, b are bfloat16')
    _require(a.is_contiguous() and b.is_contiguous() and d.is_contiguous(), 'a, b, d are contiguous')   # both K-major (gemm.hpp:541)
    _require(masked_m.is_contiguous() and masked_m.dtype == torch.int32, 'masked_m is contiguous int32')
    (g, m, k), (g_, n, k_), (g__, m_, n_) = a.shape, b.shape, d.shape
    _require(g == g_ == g__ == masked_m.numel(), 'group counts agree')
    _require(m == m_ and n == n_ and k == k_, 'm == m_ and n == n_ and k == k_')
    _require(expected_m > 0 and m > 0 and n > 0 and k > 0 and g > 0, 'positive sizes')
    _require(k % 8 == 0, 'k % 8 == 0')
    _require(d.dtype == torch.bfloat16, 'd.dtype == bfloat16')
    check(lib().dgb200_m_grouped_bf16_gemm_nt_masked(a.data_ptr(), b.data_ptr(), d.data_ptr(), masked_m.data_ptr(), g, m, n, k,
                                                     int(expected_m), _stream()))


def k_grouped_bf16_gemm_tn_contiguous(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, ks_cpu: Optional[List[int]],
                                      grouped_layout: torch.Tensor, c: Optional[torch.Tensor] = None, compiled_dims: str = 'mn',
                                      use_psum_layout: bool = False) -> None:
    """Weight gradient D[g] = C[g] + A[k_g,:M].T @ B[k_g,:N]: A [sum_k, M], B [sum_k, N] BF16 (both MN-major), D = C [G, M, N]
    FP32 accumulated in place (gemm.hpp:566-608)."""
    _require(a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, 'a, b are bfloat16')
    k_alignment = get_mk_alignment_for_contiguous_layout()
    _require(k_alignment % 32 == 0, 'k_alignment % 32 == 0')
    _require(d.dim() == 3 and a.dim() == 2 and b.dim() == 2, 'd is 3-D, a and b are 2-D')
    num_groups, m, n = d.shape
    (sum_k_a, m_), (sum_k_b, n_) = a.shape, b.shape
    _require(grouped_layout.is_contiguous() and grouped_layout.dtype == torch.int32 and grouped_layout.numel() == num_groups,
             'grouped_layout is a contiguous int32 [num_groups] tensor')
    if ks_cpu is not None and len(ks_cpu) > 0:
        _require(len(ks_cpu) == num_groups, 'len(ks_cpu) == num_groups')
        _require(all(k % k_alignment == 0 for k in ks_cpu), 'k % k_alignment == 0')
        sum_k = sum(ks_cpu)
    else:
        _require(use_psum_layout, 'ks_cpu may only be omitted with use_psum_layout')
        sum_k = sum_k_a
    _require(m == m_ and n == n_ and sum_k == sum_k_a and sum_k == sum_k_b, 'shapes agree')
    _require(a.is_contiguous() and b.is_contiguous() and d.is_contiguous(), 'a, b, d are contiguous')
    _require(c is not None and c.is_contiguous(), 'c is required and contiguous')
    _require(d.dtype == torch.float32, 'd.dtype == float')
    _require(m % 8 == 0 and n % 8 == 0, '16-byte aligned rows')
    if _early_return(m, n, sum_k, d, c):
        return
    check(lib().dgb200_k_grouped_bf16_gemm_tn_contiguous(a.data_ptr(), b.data_ptr(), d.data_ptr(), grouped_layout.data_ptr(),
                                                         num_groups, m, n, sum_k, int(use_psum_layout), _stream()))


def _bmk_bnk_mn(a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, c: Optional[torch.Tensor]) -> None:
    """D[m,n] (+)= sum_b A[b] @ B[b].T (csrc/apis/einsum.hpp:22-60). FP32 D is accumulated in place (`c` must be `d`); a BF16 D
    goes through a zeroed FP32 workspace and one cast, exactly as the reference does."""
    _require(d.dtype in (torch.float32, torch.bfloat16), 'd is float or bfloat16')
    if d.dtype == torch.bfloat16:
        _require(c is None, 'c is not supported with a BF16 output')                 # einsum.hpp:30
        ws = torch.zeros(d.shape, dtype=torch.float32, device=d.device)
        _bmk_bnk_mn(a, b, ws, ws)
        d.copy_(ws)
        return
    _require(c is not None and c.data_ptr() == d.data_ptr() and c.shape == d.shape and c.stride() == d.stride(),
             'FP32 output: c must be d (accumulated in place)')                       # einsum.hpp:26
    _require(a.is_contiguous() and b.is_contiguous() and d.is_contiguous(), 'a, b, d are contiguous')
    _require(a.dim() == 3 and b.dim() == 3 and d.dim() == 2, 'a, b are 3-D, d is 2-D')
    (s, m, k), (s_, n, k_) = a.shape, b.shape
    _require(s == s_ and k == k_ and d.shape == (m, n), 'shapes agree')
    _require(k % 64 == 0, 'k % 64 == 0')
    _require(a.data_ptr() % 16 == 0 and b.data_ptr() % 16 == 0, '16-byte aligned operands')
    check(lib().dgb200_bf16_bmk_bnk_mn(a.data_ptr(), b.data_ptr(), d.data_ptr(), s, m, n, k, _stream()))


def einsum(expr: str, a: torch.Tensor, b: torch.Tensor, d: torch.Tensor, c: Optional[torch.Tensor] = None,
           use_cublaslt: bool = False) -> None:
    """The BF16 contractions of the reference (csrc/apis/einsum.hpp:22-136): 'bhr,hdr->bhd' and 'bhd,hdr->bhr', each one batched
    GEMM over permuted views (batch = h, m = b), no copies; 'bmk,bnk->mn', the batch-reduction form (a separate kernel in the
    reference, impls/sm100_bmk_bnk_mn.cuh; here a scheduler type of the same kernel)."""
    _require(a.dtype == torch.bfloat16 and b.dtype == torch.bfloat16, 'a, b are bfloat16')
    _require(not use_cublaslt, 'use_cublaslt is not available in this library')
    if expr == 'bmk,bnk->mn':
        return _bmk_bnk_mn(a, b, d, c)
    if expr not in ('bhr,hdr->bhd', 'bhd,hdr->bhr'):
        raise RuntimeError(f'Unsupported einsum expression: {expr}')
    _require(c is None, 'c is not supported for this expression')                     # einsum.hpp:127,130
    _require(a.dim() == 3 and b.dim() == 3 and d.dim() == 3, 'a, b, d are 3-D')
    _require(d.dtype == torch.bfloat16 and a.stride(2) == 1 and b.stride(2) == 1 and d.stride(2) == 1, 'BF16, innermost contiguous')
    bsz, h, k = a.shape
    if expr == 'bhr,hdr->bhd':
        h_, n, k_ = b.shape
        major_b = _K_MAJOR
    else:
        h_, k_, n = b.shape
        major_b = _MN_MAJOR
    _require(d.shape == (bsz, h, n) and h == h_ and k == k_, 'shapes agree')
    if bsz == 0 or h == 0 or n == 0:
        return
    for t, what in ((a, 'a'), (b, 'b')):
        _require(t.data_ptr() % 16 == 0 and all(st % 8 == 0 for st in t.stride()[:2]), f'{what}: 16-byte aligned rows')
    _require(k % 8 == 0 and n % 8 == 0 if major_b == _MN_MAJOR else k % 8 == 0, '16-byte aligned rows')
    check(lib().dgb200_bf16_bmm(a.data_ptr(), b.data_ptr(), d.data_ptr(), h, bsz, n, k, a.stride(0), b.stride(1), d.stride(0),
                                a.stride(1), b.stride(0), d.stride(1), major_b, _stream()))


def _t(pair: TensorPair, d0: int = 0, d1: int = 1) -> TensorPair:
    return pair[0].transpose(d0, d1), pair[1].transpose(d0, d1)


def fp8_gemm_nn(a, b, d, c=None, recipe=None, recipe_a=None, recipe_b=None, compiled_dims='nk', disable_ue8m0_cast=False):
    """B given as [K,N] (gemm.hpp:126-137)."""
    fp8_gemm_nt(a, _t(b), d, c, recipe, recipe_a, recipe_b, compiled_dims, disable_ue8m0_cast)


def fp8_gemm_tn(a, b, d, c=None, recipe=None, recipe_a=None, recipe_b=None, compiled_dims='mn', disable_ue8m0_cast=False):
    """A given as [K,M], B as [K,N] (gemm.hpp:139-151)."""
    fp8_gemm_nt(_t(a), _t(b), d, c, recipe, recipe_a, recipe_b, compiled_dims, disable_ue8m0_cast)


def fp8_gemm_tt(a, b, d, c=None, recipe=None, recipe_a=None, recipe_b=None, compiled_dims='mn', disable_ue8m0_cast=False):
    """A given as [K,M], B as [N,K] (gemm.hpp:153-164)."""
    fp8_gemm_nt(_t(a), b, d, c, recipe, recipe_a, recipe_b, compiled_dims, disable_ue8m0_cast)


# ------------------------------------------------------------------------------------------------ M-grouped
def m_grouped_fp8_gemm_nt_contiguous(a: TensorPair, b: TensorPair, d: torch.Tensor, grouped_layout: torch.Tensor,
                                     recipe=None, recipe_a=None, recipe_b=None, compiled_dims: str = 'nk',
                                     disable_ue8m0_cast: bool = False, use_psum_layout: bool = False,
                                     ensure_zero_padding: bool = True,
                                     expected_m_for_psum_layout: Optional[int] = None) -> None:
    """A [M_sum,K] rows grouped by expert, B [G,N,K], D [M_sum,N] BF16 (gemm.hpp:166-232)."""
    (a_t, sfa), (b_t, sfb) = a, b
    _check_fp8(a_t), _check_fp8(b_t)
    major_a, major_b = _major_ab(a_t), _major_ab(b_t)
    _require(major_a == _K_MAJOR, 'major_a == K (m-grouped GEMMs need K-major A)')
    _require(grouped_layout.is_contiguous(), 'grouped_layout.is_contiguous()')
    _require(a_t.dim() == 2 and b_t.dim() == 3 and d.dim() == 2, 'a 2-D, b 3-D, d 2-D')
    (m, k), (num_groups, n, k_), (m_, n_) = a_t.shape, b_t.shape, d.shape
    _require(m == m_ and n == n_ and k == k_, 'm == m_ and n == n_ and k == k_')
    _require(n > 0 and k > 0 and num_groups > 0, 'n > 0 and k > 0 and num_groups > 0')
    _require(d.dtype == torch.bfloat16, 'd.dtype == bfloat16')
    _require(grouped_layout.dtype == torch.int32, 'grouped_layout.dtype == int32')
    if use_psum_layout:
        _require(grouped_layout.dim() == 1 and grouped_layout.numel() == num_groups, 'grouped_layout is [num_groups]')
    else:
        _require(grouped_layout.dim() == 1 and grouped_layout.numel() == m, 'grouped_layout is [m]')
        _require(expected_m_for_psum_layout is None, 'expected_m_for_psum_layout needs use_psum_layout')
    _check_cd(d)
    if m == 0:
        return
    sfa_t, sfb_t, gran_k_a, gran_k_b = _layout.transform_sf_pair_into_required_layout(
        sfa, sfb, m, n, k, recipe, recipe_a, recipe_b, None, num_groups, disable_ue8m0_cast,
        grouped_layout if use_psum_layout else None)
    _require(sfa_t.dtype == torch.int32 and sfb_t.dtype == torch.int32, 'Unsupported architecture or scaling factor types')
    ldb = b_t.stride(1) if major_b == _K_MAJOR else b_t.stride(2)
    check(lib().dgb200_m_grouped_fp8_gemm_nt_contiguous(
        a_t.data_ptr(), sfa_t.data_ptr(), b_t.data_ptr(), sfb_t.data_ptr(), d.data_ptr(), grouped_layout.data_ptr(),
        num_groups, m, n, k, a_t.stride(0), ldb, d.stride(0), major_b, sfa_t.stride(-1), sfb_t.stride(-1),
        gran_k_a, gran_k_b, int(use_psum_layout), int(ensure_zero_padding),
        -1 if expected_m_for_psum_layout is None else int(expected_m_for_psum_layout), _stream()))


def m_grouped_fp8_gemm_nn_contiguous(a, b, d, grouped_layout, recipe=None, recipe_a=None, recipe_b=None,
                                     compiled_dims='nk', disable_ue8m0_cast=False, use_psum_layout=False,
                                     ensure_zero_padding=True) -> None:
    """B given as [G,K,N] (gemm.hpp:234-248)."""
    m_grouped_fp8_gemm_nt_contiguous(a, _t(b, 1, 2), d, grouped_layout, recipe, recipe_a, recipe_b, compiled_dims,
                                     disable_ue8m0_cast, use_psum_layout, ensure_zero_padding, None)


def m_grouped_fp8_gemm_nt_masked(a: TensorPair, b: TensorPair, d: torch.Tensor, masked_m: torch.Tensor,
                                 expected_m: int, recipe=None, recipe_a=None, recipe_b=None,
                                 compiled_dims: str = 'nk', disable_ue8m0_cast: bool = False) -> None:
    """A [G,M_max,K], B [G,N,K], D [G,M_max,N] BF16; rows >= masked_m[g] are ignored. `masked_m` stays on the device
    (CUDA-graph safe). Reference: gemm.hpp:250-297."""
    (a_t, sfa), (b_t, sfb) = a, b
    _check_fp8(a_t), _check_fp8(b_t)
    _require(_major_ab(a_t) == _K_MAJOR and _major_ab(b_t) == _K_MAJOR, 'major_a == K and major_b == K')
    _require(masked_m.is_contiguous(), 'masked_m.is_contiguous()')
    _require(a_t.dim() == 3 and b_t.dim() == 3 and d.dim() == 3, 'a, b, d are 3-D')
    (g, m, k), (g_, n, k_), (g__, m_, n_) = a_t.shape, b_t.shape, d.shape
    _require(g == g_ == g__ == masked_m.numel(), 'group counts agree')
    _require(m == m_ and n == n_ and k == k_, 'm == m_ and n == n_ and k == k_')
    _require(expected_m > 0 and m > 0 and n > 0 and k > 0 and g > 0, 'positive sizes')
    _require(d.dtype == torch.bfloat16, 'd.dtype == bfloat16')
    _require(masked_m.dtype == torch.int32, 'masked_m.dtype == int32')
    _check_cd(d)
    _require(d.stride(1) == n, 'd is densely batched')
    sfa_t, sfb_t, gran_k_a, gran_k_b = _layout.transform_sf_pair_into_required_layout(
        sfa, sfb, m, n, k, recipe, recipe_a, recipe_b, g, g, disable_ue8m0_cast)
    _require(sfa_t.dtype == torch.int32 and sfb_t.dtype == torch.int32, 'Unsupported architecture or scaling factor types')
    check(lib().dgb200_m_grouped_fp8_gemm_nt_masked(
        a_t.data_ptr(), sfa_t.data_ptr(), b_t.data_ptr(), sfb_t.data_ptr(), d.data_ptr(), masked_m.data_ptr(),
        g, m, n, k, int(expected_m), sfa_t.stride(-1), sfb_t.stride(-1), gran_k_a, gran_k_b, _stream()))


# ------------------------------------------------------------------------------------------------ K-grouped
def k_grouped_fp8_gemm_tn_contiguous(a: TensorPair, b: TensorPair, d: torch.Tensor, ks_cpu: Optional[List[int]],
                                     grouped_layout: torch.Tensor, c: Optional[torch.Tensor] = None,
                                     recipe: Tuple[int, int, int] = (1, 1, 128), compiled_dims: str = 'mn',
                                     use_psum_layout: bool = False) -> None:
    """Weight gradient D[g] = C[g] + A[k_g,:M].T @ B[k_g,:N]: A [sum_k, M], B [sum_k, N] FP8 (both MN-major), D = C
    [G, M, N] FP32 accumulated in place, per-(gran_k x 1) scale factors (gemm.hpp:299-346)."""
    (a_t, sfa), (b_t, sfb) = a, b
    _check_fp8(a_t), _check_fp8(b_t)
    _require(recipe[0] == 1 and recipe[1] == 1, 'recipe is (1, 1, gran_k)')
    gran_k = recipe[2]
    _require(gran_k in (32, 128), 'gran_k == 32 or gran_k == 128')
    k_alignment = get_mk_alignment_for_contiguous_layout()
    _require(k_alignment % 32 == 0, 'k_alignment % 32 == 0')
    _require(d.dim() == 3 and a_t.dim() == 2 and b_t.dim() == 2, 'd is 3-D, a and b are 2-D')
    num_groups, m, n = d.shape
    (sum_k_a, m_), (sum_k_b, n_) = a_t.shape, b_t.shape
    _require(grouped_layout.is_contiguous() and grouped_layout.dtype == torch.int32 and grouped_layout.numel() == num_groups,
             'grouped_layout is a contiguous int32 [num_groups] tensor')
    if ks_cpu is not None and len(ks_cpu) > 0:
        _require(len(ks_cpu) == num_groups, 'len(ks_cpu) == num_groups')
        _require(all(k % k_alignment == 0 for k in ks_cpu), 'k % k_alignment == 0')
        sum_k = sum(ks_cpu)
    else:
        _require(use_psum_layout, 'ks_cpu may only be omitted with use_psum_layout')
        sum_k = sum_k_a
    _require(m == m_ and n == n_ and sum_k == sum_k_a and sum_k == sum_k_b, 'shapes agree')
    _require(a_t.is_contiguous() and b_t.is_contiguous() and d.is_contiguous(), 'a, b, d are contiguous')
    _require(c is not None and c.is_contiguous(), 'c is required and contiguous')
    _require(d.dtype == torch.float32, 'd.dtype == float')
    if _early_return(m, n, sum_k, d, c):
        return
    sfa_t = _layout.transform_k_grouped_sf_into_required_layout(sfa, ks_cpu, grouped_layout, recipe, k_alignment, use_psum_layout)
    sfb_t = _layout.transform_k_grouped_sf_into_required_layout(sfb, ks_cpu, grouped_layout, recipe, k_alignment, use_psum_layout)
    _require(sfa_t.size(0) == sfb_t.size(0), 'sfa and sfb have the same number of packed rows')
    check(lib().dgb200_k_grouped_fp8_gemm_tn_contiguous(
        a_t.data_ptr(), sfa_t.data_ptr(), b_t.data_ptr(), sfb_t.data_ptr(), d.data_ptr(), grouped_layout.data_ptr(),
        num_groups, m, n, sum_k, sfa_t.size(0), gran_k, int(use_psum_layout), _stream()))


def k_grouped_fp8_gemm_nt_contiguous(a, b, d, ks_cpu, grouped_layout, c=None, recipe=(1, 1, 128), compiled_dims='mn',
                                     use_psum_layout=False) -> None:
    """SM90-only in the reference (gemm.hpp:393-399 -> 'Unsupported architecture' on SM100)."""
    raise RuntimeError('Unsupported architecture')
