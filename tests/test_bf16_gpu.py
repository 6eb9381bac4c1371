"""GPU parity tests of the BF16-operand GEMMs (fp8_gemm_kernel<..., kBf16AB>; reference bf16_gemm_nt,
m_grouped_bf16_gemm_nt_contiguous, m_grouped_bf16_gemm_nt_masked, csrc/apis/gemm.hpp:404-564).

Floating-point kernel, so the checker is a plain PyTorch matmul of the same op: BF16 inputs are exact in FP64, and
`a.double() @ b.double().T` differs from the kernel only by the FP32 accumulation order inside the tensor core (tolerance in
`_close`; the reference's own bound for its BF16 kernels is calc_diff < 1e-5, tests/test_bf16.py). Against the reference's
kernel on the same inputs the output is bit-identical (digests in tests/golden/gpu_digests.json, keys `bf16_*`;
tools/bf16_bench.py re-checks it live, MN-major and k-grouped forms included)."""
import json
import os
import random
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, 'golden'))


@pytest.fixture(scope='module')
def dg():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    import deepgemm_b200
    from deepgemm_b200 import _lib
    _lib.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    return deepgemm_b200


@pytest.fixture
def no_split_k(dg):
    """Cluster split-K (dense, K-major, M <= ~256) adds the K slices in slice order: within FP32 rounding of the one-pass
    kernel, not bit-identical to it. Bit-for-bit comparisons switch it off, as `set_split_k(False)` does for a user."""
    dg.set_split_k(False)
    yield
    dg.set_split_k(True)


def _close(d, ref, what=''):
    """`ref` is the FP64 product of the (exact) BF16 inputs. Stated tolerance: FP32 outputs within 3e-5 x max|D| (FP32
    accumulation of up to 7168 terms in the tensor core's own order), BF16 outputs within one BF16 rounding step of it."""
    from deepgemm_b200.testing import calc_diff
    ref = ref.double()
    assert not torch.isnan(d.float()).any(), what
    assert calc_diff(d, ref) < 1e-5, what
    scale = float(ref.abs().max().clamp(min=1.0))
    if d.dtype == torch.float32:
        assert float((d.double() - ref).abs().max()) < 3e-5 * scale, what
    else:
        err = (d.double() - ref).abs()
        assert bool((err <= ref.abs() * 2.0 ** -7 + 3e-5 * scale).all()), what


@pytest.mark.parametrize('m,n,k', [(128, 128, 128), (1, 576, 512), (64, 4096, 7168), (300, 2112, 1536), (4096, 4096, 2048), (97, 136, 200)])
@pytest.mark.parametrize('out_dtype', [torch.bfloat16, torch.float32])
def test_bf16_gemm_nt_matches_fp32_matmul(dg, m, n, k, out_dtype):
    gen = torch.Generator(device='cuda').manual_seed(m + n + k)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    buf = torch.full((m + 8, n + 16), 222.0, device='cuda', dtype=out_dtype)
    d = buf[:m, :n]
    dg.bf16_gemm_nt(a, b, d)
    ref = a.double() @ b.double().t()
    _close(d, ref, f'{m}x{n}x{k}')
    assert bool((buf[m:] == 222.0).all()) and bool((buf[:, n:] == 222.0).all())
    # accumulate into C
    c = (torch.randn((m, n), device='cuda') * 4).to(out_dtype)
    d2 = c.clone()
    dg.bf16_gemm_nt(a, b, d2, c=d2)
    want = (ref.to(torch.bfloat16).float() + c.float()) if out_dtype == torch.bfloat16 else ref.float() + c
    err = (d2.float() - want).abs()
    mag = ref.float().abs() + c.float().abs()
    assert bool((err <= mag * 2.0 ** -6 + 1e-5 * mag.max()).all())


@pytest.mark.parametrize('m,n,k', [(64, 4096, 7168), (128, 2112, 7168), (17, 576, 4096), (192, 4096, 7168), (256, 2112, 7168)])
def test_bf16_cluster_split_k_stays_within_fp32_rounding(dg, m, n, k):
    """Default configuration at small / medium M: K slices reduced through distributed shared memory (as for FP8 operands)."""
    from deepgemm_b200 import _lib
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16)
    ref = a.double() @ b.double().t()
    for out_dtype in (torch.bfloat16, torch.float32):
        d = torch.empty((m, n), device='cuda', dtype=out_dtype)
        dg.bf16_gemm_nt(a, b, d)
        cfg = _lib.last_config()
        _close(d, ref, f'{m}x{n}x{k} {cfg}')
        dg.set_split_k(False)
        try:
            one = torch.empty_like(d)
            dg.bf16_gemm_nt(a, b, one)
        finally:
            dg.set_split_k(True)
        if cfg['cluster_split']:
            frac = float((one != d).float().mean())
            assert frac < (0.01 if out_dtype == torch.bfloat16 else 1.0), frac      # BF16: a rounding step on a few outputs at most
        else:
            assert torch.equal(one, d)
        c = torch.randn((m, n), device='cuda').to(out_dtype)
        acc = c.clone()
        dg.bf16_gemm_nt(a, b, acc, c=acc)
        want = ref.to(torch.bfloat16).double() + c.double() if out_dtype == torch.bfloat16 else ref + c.double()
        err = (acc.double() - want).abs()
        mag = ref.abs() + c.double().abs()
        assert bool((err <= mag * 2.0 ** -6 + 3e-5 * float(mag.max())).all())


def test_bf16_gemm_every_tile_config_gives_identical_bits(dg, monkeypatch):
    m, n, k = 500, 1024, 1024
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16)
    base = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    dg.bf16_gemm_nt(a, b, base)
    for bm, st in ((16, 3), (64, 0), (128, 0), (240, 0), (208, 2)):
        monkeypatch.setenv('DGB200_BLOCK_M', str(bm))
        if st:
            monkeypatch.setenv('DGB200_STAGES', str(st))
        d = torch.empty_like(base)
        dg.bf16_gemm_nt(a, b, d)
        assert torch.equal(d, base), (bm, st)
        monkeypatch.delenv('DGB200_STAGES', raising=False)


def test_bf16_transposed_wrappers_on_k_major_views(dg, no_split_k):
    a = torch.randn((256, 512), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((384, 512), device='cuda', dtype=torch.bfloat16)
    d0 = torch.empty((256, 384), device='cuda', dtype=torch.bfloat16)
    dg.bf16_gemm_nt(a, b, d0)
    d1 = torch.empty_like(d0)
    dg.bf16_gemm_nn(a, b.t(), d1)                               # B given as the [K, N] view of a K-major tensor
    assert torch.equal(d0, d1)
    d2 = torch.empty_like(d0)
    dg.bf16_gemm_tt(a.t(), b, d2)
    assert torch.equal(d0, d2)


@pytest.mark.parametrize('m,n,k', [(128, 128, 128), (64, 4096, 7168), (304, 2112, 1536), (4096, 4096, 2048), (96, 136, 200), (8, 576, 512)])
@pytest.mark.parametrize('majors', ['nn', 'tn', 'tt'])
def test_bf16_gemm_mn_major_operands_give_the_k_major_bits(dg, no_split_k, m, n, k, majors):
    """Genuinely MN-major operands (bf16_gemm_{nn,tn,tt}, gemm.hpp:440-462): the same products summed in the same order, so the
    output must equal the K-major launch bit for bit."""
    gen = torch.Generator(device='cuda').manual_seed(m * 3 + n + k)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    a_km = a.t().contiguous()      # [K, M]
    b_kn = b.t().contiguous()      # [K, N]
    for out_dtype in (torch.bfloat16, torch.float32):
        base = torch.empty((m, n), device='cuda', dtype=out_dtype)
        dg.bf16_gemm_nt(a, b, base)
        _close(base, a.double() @ b.double().t())
        buf = torch.full((m + 8, n + 16), 222.0, device='cuda', dtype=out_dtype)
        d = buf[:m, :n]
        if majors == 'nn':
            dg.bf16_gemm_nn(a, b_kn, d)
        elif majors == 'tn':
            dg.bf16_gemm_tn(a_km, b_kn, d)
        else:
            dg.bf16_gemm_tt(a_km, b, d)
        assert torch.equal(d, base), (majors, out_dtype)
        assert bool((buf[m:] == 222.0).all()) and bool((buf[:, n:] == 222.0).all())
    c = torch.randn((m, n), device='cuda')
    want = c.clone()
    dg.bf16_gemm_nt(a, b, want, c=want)
    got = c.clone()
    dg.bf16_gemm_tn(a_km, b_kn, got, c=got)
    assert torch.equal(got, want)


def test_bf16_mn_major_every_tile_height_gives_identical_bits(dg, monkeypatch):
    m, n, k = 512, 640, 1088
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16)
    a_km, b_kn = a.t().contiguous(), b.t().contiguous()
    base = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    dg.bf16_gemm_nt(a, b, base)
    for bm in (32, 64, 96, 128, 160, 192, 224):            # token swizzle atoms of 32 / 64 / 128 bytes
        monkeypatch.setenv('DGB200_BLOCK_M', str(bm))
        d = torch.empty_like(base)
        dg.bf16_gemm_tn(a_km, b_kn, d)
        assert torch.equal(d, base), bm


@pytest.mark.parametrize('use_psum', [False, True])
def test_m_grouped_bf16_nn_contiguous_gives_the_nt_bits(dg, use_psum):
    g, n, k, alignment = 4, 512, 768, 128
    ms = [100, 0, 256, 77]
    aligned = [(x + alignment - 1) // alignment * alignment for x in ms]
    m = sum(aligned)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16)
    layout = torch.empty(g if use_psum else m, device='cuda', dtype=torch.int32)
    s = 0
    for i, (mi, ai) in enumerate(zip(ms, aligned)):
        if use_psum:
            layout[i] = s + mi
        else:
            layout[s:s + mi] = i
            layout[s + mi:s + ai] = -1
        s += ai
    d0 = torch.zeros((m, n), device='cuda', dtype=torch.bfloat16)
    d1 = torch.zeros_like(d0)
    dg.m_grouped_bf16_gemm_nt_contiguous(a, b, d0, layout, use_psum_layout=use_psum)
    dg.m_grouped_bf16_gemm_nn_contiguous(a, b.transpose(1, 2).contiguous(), d1, layout, use_psum_layout=use_psum)
    assert torch.equal(d0, d1)


@pytest.mark.parametrize('use_psum', [False, True])
def test_k_grouped_bf16_gemm_tn_contiguous(dg, use_psum):
    """Weight gradient D[g] = C[g] + A_g^T B_g over K segments (gemm.hpp:566-608; tests/test_bf16.py k-grouped case)."""
    random.seed(11 + use_psum)
    g, m, n, k_alignment = 5, 384, 264, 128
    ks = [k_alignment * random.randint(1, 6) for _ in range(g)]
    ks[1] = 0
    if use_psum:
        ks_real = [max(0, kk - random.randint(0, 60)) if kk else 0 for kk in ks]      # unaligned ends, aligned starts
    else:
        ks_real = ks
    sum_k = sum(ks)
    a = torch.randn((sum_k, m), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((sum_k, n), device='cuda', dtype=torch.bfloat16)
    c = torch.randn((g, m, n), device='cuda', dtype=torch.float32)
    ref = c.clone()
    s = 0
    ends = []
    for i, (kk, kr) in enumerate(zip(ks, ks_real)):
        a[s + kr:s + kk] = 0                     # rows between a group's end and the next aligned start are zero padding
        b[s + kr:s + kk] = 0
        ref[i] += a[s:s + kr].float().t() @ b[s:s + kr].float()
        ends.append(s + kr)
        s += kk
    if use_psum:
        layout = torch.tensor(ends, device='cuda', dtype=torch.int32)
        d = c.clone()
        dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d, None, layout, c=d, use_psum_layout=True)
    else:
        layout = torch.tensor(ks, device='cuda', dtype=torch.int32)
        d = c.clone()
        dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d, ks, layout, c=d)
    _close(d, ref, f'k-grouped psum={use_psum}')


@pytest.mark.parametrize('h,r,dd', [(16, 512, 128), (8, 1024, 256)])
@pytest.mark.parametrize('bsz', [4, 100, 1024])
def test_bf16_einsum(dg, h, r, dd, bsz):
    """einsum('bhr,hdr->bhd') and ('bhd,hdr->bhr') on sliced weights, as tests/test_einsum.py:40-80 builds them."""
    fy = torch.randn((h, dd, r + 128), device='cuda', dtype=torch.bfloat16)
    y = fy[:, :, :r]
    x = torch.randn((bsz, h, r), device='cuda', dtype=torch.bfloat16)
    z = torch.empty((bsz, h, dd), device='cuda', dtype=torch.bfloat16)
    dg.einsum('bhr,hdr->bhd', x, y, z)
    _close(z, torch.einsum('bhr,hdr->bhd', x.float(), y.float()), 'bhr,hdr->bhd')
    x2 = torch.randn((bsz, h, dd), device='cuda', dtype=torch.bfloat16)
    z2 = torch.empty((bsz, h, r), device='cuda', dtype=torch.bfloat16)
    dg.einsum('bhd,hdr->bhr', x2, y, z2)
    _close(z2, torch.einsum('bhd,hdr->bhr', x2.float(), y.float()), 'bhd,hdr->bhr')
    with pytest.raises(RuntimeError):
        dg.einsum('bmk,bhk->mh', x, x, z)


@pytest.mark.parametrize('s', [1, 129, 4096])
@pytest.mark.parametrize('m,n,k', [(128, 384, 128), (256, 256, 256), (384, 128, 384), (72, 200, 64)])
def test_bf16_einsum_bmk_bnk_mn(dg, s, m, n, k):
    """The batch-reduction form, tests/test_einsum.py:16-35: FP32 D accumulated in place, BF16 D through an FP32 workspace.
    Tolerance: the reference's own (calc_diff < 1e-5); partial tiles are added with memory-side FP32 adds in no fixed order."""
    from deepgemm_b200.testing import calc_diff
    gen = torch.Generator(device='cuda').manual_seed(s + m + k)
    a = torch.randn((s, m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((s, n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    prod = torch.bmm(a.double(), b.double().mT).sum(0)
    for dtype in (torch.float32, torch.bfloat16):
        d = torch.randn((m, n), device='cuda', dtype=dtype, generator=gen)
        c = d if dtype == torch.float32 else None
        ref = (d.double() if dtype == torch.float32 else 0) + prod
        dg.einsum('bmk,bnk->mn', a, b, d, c=c)
        assert calc_diff(d, ref) < 1e-5, (s, m, n, k, dtype)
        scale = float(ref.abs().max().clamp(min=1.0))
        # up to 1.5 M products per output, accumulated in FP32 in chunks: the error grows with the length of the sum
        tol = 2e-4 * scale if dtype == torch.float32 else 2.0 ** -7 * scale
        assert float((d.double() - ref).abs().max()) <= tol
    with pytest.raises(RuntimeError):
        dg.einsum('bmk,bnk->mn', a, b, torch.empty((m, n), device='cuda'), c=None)          # FP32 needs c is d


@pytest.mark.parametrize('use_psum', [False, True])
def test_m_grouped_bf16_contiguous(dg, use_psum):
    random.seed(3 + use_psum)
    g, n, k, alignment = 6, 768, 1024, 128
    ms = [int(150 * random.uniform(0.3, 1.7)) for _ in range(g)]
    ms[2] = 0
    aligned = [(x + alignment - 1) // alignment * alignment for x in ms]
    m = sum(aligned)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16)
    layout = torch.empty(g if use_psum else m, device='cuda', dtype=torch.int32)
    valid = torch.zeros(m, dtype=torch.bool, device='cuda')
    expert = torch.zeros(m, dtype=torch.long, device='cuda')
    s = 0
    for i, (mi, ai) in enumerate(zip(ms, aligned)):
        if use_psum:
            layout[i] = s + mi
        else:
            layout[s:s + mi] = i
            layout[s + mi:s + ai] = -1
        valid[s:s + mi] = True
        expert[s:s + ai] = i
        s += ai
    d = torch.full((m, n), float('nan'), device='cuda', dtype=torch.bfloat16)
    dg.m_grouped_bf16_gemm_nt_contiguous(a, b, d, layout, use_psum_layout=use_psum)
    ref = torch.einsum('mk,mnk->mn', a.float(), b.float()[expert])
    _close(d[valid], ref[valid], f'contiguous psum={use_psum}')
    if use_psum:
        assert bool((d[~valid] == 0).all())


def test_m_grouped_bf16_masked(dg):
    g, m_max, n, k = 8, 192, 512, 1024
    a = torch.randn((g, m_max, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16)
    masked_m = torch.tensor([5, 192, 0, 64, 100, 17, 128, 33], device='cuda', dtype=torch.int32)
    d = torch.full((g, m_max, n), 1234.0, device='cuda', dtype=torch.bfloat16)
    dg.m_grouped_bf16_gemm_nt_masked(a, b, d, masked_m, 64)
    ref = torch.einsum('gmk,gnk->gmn', a.float(), b.float())
    for gi, mg in enumerate(masked_m.tolist()):
        if mg:
            _close(d[gi, :mg], ref[gi, :mg], f'masked {gi}')
        assert bool((d[gi, mg:] == 1234.0).all())


@pytest.mark.parametrize('m,n,k', [(128, 2112, 7168), (4096, 7168, 2048), (64, 576, 7168)])
def test_bf16_matches_the_reference_kernel_bit_for_bit(dg, no_split_k, m, n, k):
    path = os.path.join(HERE, 'golden', 'gpu_digests.json')
    key = f'bf16_{m}x{n}x{k}'
    digests = json.load(open(path)) if os.path.exists(path) else {}
    if key not in digests:
        pytest.skip(f'no digest for {key}')
    import cases
    a, b = cases.make_bf16(m, n, k)
    d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    dg.bf16_gemm_nt(a, b, d)
    torch.cuda.synchronize()
    assert cases.digest(d) == digests[key]
