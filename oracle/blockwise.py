"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the FP8 blockwise-scaled GEMM path (torch CPU, no GPU, no CUDA library).

What is restated, and from where (paths in the reference tree):
  * wire format of packed UE8M0 scale factors ........ tests/test_layout.py:20-42 (the reference's own bit-exact
    torch restatement), csrc/jit_kernels/impls/smxx_layout.hpp:155-178, deep_gemm/include/deep_gemm/impls/smxx_layout.cuh:116-143
  * 128x128 weight scale -> per-row broadcast ........ csrc/apis/layout.hpp:48-54
  * arithmetic of the SM100 kernel ................... impls/sm100_fp8_fp4_gemm_1d1d.cuh:374-396 (`tcgen05.mma
    kind::mxf8f6f4.block_scale`): every e4m3*e4m3 product is exact, the per-K-granule scale 2^(ea-127) * 2^(eb-127) is
    exact, products are accumulated in FP32, and the result is rounded once to BF16 (common/math.cuh:72-76).
    => dequantising both operands to FP32 (exact: 4 significant bits times a power of two) and accumulating in
       FP32/FP64 is an exact-input oracle; only the FP32 accumulation ORDER inside the tensor core is unknown, which
       bounds the difference to FP32 rounding of the accumulator (<= 1 BF16 ulp after the final cast).
  * C/D, empty-problem semantics ..................... csrc/apis/gemm.hpp:19-46
  * group -> rows mapping ............................ scheduler/gemm.cuh:160-165 (contiguous), :200-216 (masked),
    :217-237 (psum)

Pinning: `tests/golden/` holds vectors generated (a) here from the reference's pure-python helpers
(tests/golden/make_golden_cpu.py) and (b) on a B200 from the reference's own SM100 kernel
(tests/golden/make_golden_gpu.py); tests/test_oracle.py checks this file against both.
"""
from typing import List, Optional, Sequence, Tuple

import torch


def ceil_div(a: int, b: int) -> int:
    return -(-a // b)


def align(a: int, b: int) -> int:
    return ceil_div(a, b) * b


# ------------------------------------------------------------------------------------------------ SF wire format
def pack_sf_ue8m0_mn_major(sf: torch.Tensor) -> torch.Tensor:
    """FP32 power-of-two SFs [.., mn, sf_k] -> int32 [.., mn, ceil(sf_k/4)] with strides (.., 1, align(mn, 4)).
    Restates tests/test_layout.py:20-42 step by step."""
    assert sf.dtype == torch.float32 and sf.dim() in (2, 3)
    exps = (sf.contiguous().view(torch.int32) >> 23).to(torch.uint8)              # exponent byte only
    squeeze = sf.dim() == 2
    if squeeze:
        exps = exps.unsqueeze(0)
    b, mn, k = exps.shape
    amn, ak = align(mn, 4), align(k, 4)
    padded = torch.zeros((b, amn, ak), dtype=torch.uint8)
    padded[:, :mn, :k] = exps
    words = padded.view(-1).view(torch.int32).view(b, amn, ak // 4)              # 4 consecutive K bytes -> 1 word (LE)
    out = torch.zeros((b, ak // 4, amn), dtype=torch.int32).transpose(1, 2)      # MN-major storage
    out[:, :, :] = words
    out = out[:, :mn, :]
    return out.squeeze(0) if squeeze else out


def unpack_sf_ue8m0(packed: torch.Tensor, sf_k: int) -> torch.Tensor:
    """Inverse of the packing: int32 [.., mn, kp] (any strides) -> FP32 [.., mn, sf_k]."""
    assert packed.dtype == torch.int32
    dense = torch.empty(packed.shape, dtype=torch.int32).copy_(packed)            # force row-major storage
    b = dense.view(torch.uint8).view(*packed.shape, 4)                            # little endian: byte j = granule 4kp+j
    exps = b.reshape(*packed.shape[:-1], packed.shape[-1] * 4)[..., :sf_k].to(torch.int32)
    return (exps << 23).view(torch.float32)


def expand_sf(sf: torch.Tensor, mn: int, k: int, gran_mn: int, gran_k: int) -> torch.Tensor:
    """FP32 SFs [ceil(mn/gran_mn), ceil(k/gran_k)] -> per-element scale [mn, k] (csrc/apis/layout.hpp:48-54)."""
    rows = torch.arange(mn) // gran_mn
    cols = torch.arange(k) // gran_k
    return sf[rows][:, cols]


def dequant(x_fp8: torch.Tensor, sf: torch.Tensor, gran_mn: int, gran_k: int) -> torch.Tensor:
    """Exact FP32 value of every operand element: e4m3 * 2^e."""
    mn, k = x_fp8.shape
    if sf.dtype == torch.int32:
        sf = unpack_sf_ue8m0(sf, ceil_div(k, gran_k))
        gran_mn = 1
    return x_fp8.float() * expand_sf(sf.float(), mn, k, gran_mn, gran_k)


def _matmul_nt(a: torch.Tensor, b: torch.Tensor, high_precision: bool) -> torch.Tensor:
    if high_precision:
        return (a.double() @ b.double().t()).float()
    return a @ b.t()


def _finish(acc: torch.Tensor, c: Optional[torch.Tensor], out_dtype: torch.dtype) -> torch.Tensor:
    """Output rounding + accumulation semantics of the SM100 epilogue (epilogue/sm100_store_cd.cuh:112-128):
    FP32 D: acc + C in FP32.  BF16 D: acc rounded to BF16, then ONE BF16 add with C (memory-side reduce-add)."""
    if out_dtype == torch.float32:
        return acc if c is None else acc + c.float()
    r = acc.to(torch.bfloat16)
    return r if c is None else (r.float() + c.float()).to(torch.bfloat16)


# ------------------------------------------------------------------------------------------------ GEMMs
def fp8_gemm_nt(a: Tuple[torch.Tensor, torch.Tensor], b: Tuple[torch.Tensor, torch.Tensor], out_dtype=torch.bfloat16,
                c: Optional[torch.Tensor] = None, recipe: Tuple[int, int, int] = (1, 128, 128),
                high_precision: bool = True) -> torch.Tensor:
    """Oracle for fp8_fp4_gemm_nt (csrc/apis/gemm.hpp:73-124). a=(A[M,K] e4m3, SFA), b=(B[N,K] e4m3, SFB);
    SFs either FP32 (granularity from `recipe`) or packed int32 (per-row)."""
    (a_t, sfa), (b_t, sfb) = a, b
    m, k = a_t.shape
    n = b_t.shape[0]
    if m == 0 or n == 0:
        return torch.empty((m, n), dtype=out_dtype)
    if k == 0:
        return c.clone() if c is not None else torch.zeros((m, n), dtype=out_dtype)
    ad = dequant(a_t, sfa, recipe[0], recipe[2])
    bd = dequant(b_t, sfb, recipe[1], recipe[2])
    return _finish(_matmul_nt(ad, bd, high_precision), c, out_dtype)


def m_grouped_fp8_gemm_nt_contiguous(a, b, grouped_layout: torch.Tensor, recipe=(1, 128, 128), use_psum_layout=False,
                                     alignment: int = 128, high_precision: bool = True) -> Tuple[torch.Tensor, torch.Tensor]:
    """Oracle for m_grouped_fp8_fp4_gemm_nt_contiguous (gemm.hpp:166-232). Returns (D [M,N] bf16, valid-row mask).
    Non-psum: row r uses expert grouped_layout[r]; padding rows (-1) are computed against expert 0
    (scheduler/gemm.cuh:161) -- they are reported as invalid in the mask. Psum: expert g owns rows
    [align(end_{g-1}, alignment), end_g)."""
    (a_t, sfa), (b_t, sfb) = a, b
    m, k = a_t.shape
    g, n, _ = b_t.shape
    ad = dequant(a_t, sfa, recipe[0], recipe[2])
    d = torch.zeros((m, n), dtype=torch.bfloat16)
    valid = torch.zeros(m, dtype=torch.bool)
    layout = grouped_layout.tolist()
    if use_psum_layout:
        start = 0
        for gi, end in enumerate(layout):
            if end > start:
                bd = dequant(b_t[gi], sfb[gi], recipe[1], recipe[2])
                d[start:end] = _matmul_nt(ad[start:end], bd, high_precision).to(torch.bfloat16)
                valid[start:end] = True
            start = align(end, alignment)
    else:
        ids = torch.tensor(layout)
        for gi in range(g):
            rows = (ids == gi).nonzero().flatten()
            if rows.numel():
                bd = dequant(b_t[gi], sfb[gi], recipe[1], recipe[2])
                d[rows] = _matmul_nt(ad[rows], bd, high_precision).to(torch.bfloat16)
                valid[rows] = True
    return d, valid


def m_grouped_fp8_gemm_nt_masked(a, b, masked_m: torch.Tensor, recipe=(1, 128, 128),
                                 high_precision: bool = True) -> torch.Tensor:
    """Oracle for m_grouped_fp8_fp4_gemm_nt_masked (gemm.hpp:250-297): D [G,M_max,N] bf16, rows >= masked_m[g] zero
    (the kernel leaves them untouched; tests compare valid rows only, tests/test_fp8_fp4.py:166-174)."""
    (a_t, sfa), (b_t, sfb) = a, b
    g, m_max, k = a_t.shape
    n = b_t.shape[1]
    d = torch.zeros((g, m_max, n), dtype=torch.bfloat16)
    for gi, mg in enumerate(masked_m.tolist()):
        if mg > 0:
            ad = dequant(a_t[gi, :mg], sfa[gi, :mg], recipe[0], recipe[2])
            bd = dequant(b_t[gi], sfb[gi], recipe[1], recipe[2])
            d[gi, :mg] = _matmul_nt(ad, bd, high_precision).to(torch.bfloat16)
    return d


def k_grouped_fp8_gemm_tn_contiguous(a, b, c: torch.Tensor, ks: Sequence[int], gran_k: int = 128,
                                     group_ends: Optional[Sequence[int]] = None, high_precision: bool = True) -> torch.Tensor:
    """Oracle for k_grouped_fp8_gemm_tn_contiguous (gemm.hpp:299-346): a=(A [sum_k, M] e4m3, SFA [sum ceil(k_g/gran_k), M]),
    b likewise with N; D[g] = C[g] + A_g^T B_g in FP32. `group_ends` (psum layout): group g occupies rows
    [end_g - k_g, end_g) of A/B; default: groups back to back. SF rows are compact per group (ceil(k_g/gran_k) each)."""
    (a_t, sfa), (b_t, sfb) = a, b
    d = c.clone().float()
    if group_ends is None:
        group_ends, acc = [], 0
        for kg in ks:
            acc += kg
            group_ends.append(acc)
    sf_row = 0
    for gi, (kg, end) in enumerate(zip(ks, group_ends)):
        if kg == 0:
            continue
        start = end - kg
        n_sf = ceil_div(kg, gran_k)
        rows = torch.arange(kg) // gran_k
        ad = a_t[start:end].float() * sfa[sf_row:sf_row + n_sf][rows]
        bd = b_t[start:end].float() * sfb[sf_row:sf_row + n_sf][rows]
        prod = (ad.double().t() @ bd.double()).float() if high_precision else ad.t() @ bd
        d[gi] += prod
        sf_row += n_sf
    return d


def pack_sf_ue8m0_k_grouped(sf: torch.Tensor, ks: Sequence[int], gran_k: int) -> torch.Tensor:
    """K-grouped packed SFs: each group's granules padded to a multiple of 4, packed 4 per int32 along K
    (restates the per-group use of the wire format in tests/test_layout.py:93-98)."""
    outs, row = [], 0
    for kg in ks:
        n_sf = ceil_div(kg, gran_k)
        if n_sf == 0:
            continue
        part = sf[row:row + n_sf].t().contiguous()                       # [mn, n_sf]
        outs.append(torch.empty((part.shape[0], ceil_div(n_sf, 4)), dtype=torch.int32).copy_(pack_sf_ue8m0_mn_major(part)).t())
        row += n_sf
    return torch.cat(outs) if outs else torch.empty((0, sf.shape[1]), dtype=torch.int32)


# ------------------------------------------------------------------------------------------------ CPU baseline
def bf16_emulated_gemm_nt(a, b, recipe=(1, 128, 128)) -> torch.Tensor:
    """The CPU baseline BASELINE.md section 4 names: dequantise FP8 x UE8M0 to BF16 (exact -- 4 significant bits and
    an in-range exponent) and run torch.matmul in BF16 with FP32 accumulation on the host cores."""
    (a_t, sfa), (b_t, sfb) = a, b
    ad = dequant(a_t, sfa, recipe[0], recipe[2]).to(torch.bfloat16)
    bd = dequant(b_t, sfb, recipe[1], recipe[2]).to(torch.bfloat16)
    return ad @ bd.t()
