"""GPU parity tests, round 2 (run with `-m gpu` on a B200): the TMA-store epilogue, fp8_gemm_nt_skip_head_mid, fp8_einsum /
fp8_bmm, the CUDA activation quantiser, gran_k = 32 and mixed recipes, the reference's `enumerate_normal` shape list and
BASELINE configs 3 / 4 at full size against SHA-256 digests of the reference kernel's own output
(tests/golden/gpu_digests.json, generated on a B200 by tests/golden/make_golden_digests.py).

Tolerances are those of tests/test_gemm_gpu.py: bit-exact versus the reference kernel and between our own kernel
variants (same instruction, same K order); FP32-accumulation-order noise versus the FP64-accumulated CPU oracle.
"""
import json
import os
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
    _lib.lib()  # fail loudly if the CUDA library is missing: there is no fallback
    return deepgemm_b200


@pytest.fixture(scope='module')
def digests():
    path = os.path.join(HERE, 'golden', 'gpu_digests.json')
    if not os.path.exists(path):
        pytest.skip('gpu_digests.json not generated yet')
    with open(path) as f:
        return json.load(f)


def _quant_dense(m, n, k, seed=0):
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    g = torch.Generator(device='cuda').manual_seed(seed)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=g)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=g)
    return a, b, per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True)


def _cpu(pair):
    return pair[0].cpu(), pair[1].cpu()


# ------------------------------------------------------------------------------------------------ TMA-store epilogue
@pytest.mark.parametrize('m,n,k', [(4096, 4096, 7168), (500, 1024, 2048), (2000, 1000, 512), (240, 136, 384), (4096, 7168, 2048),
                                   (97, 384, 256), (130, 64, 128)])
def test_tma_store_epilogue_is_bit_identical_to_direct_stores(dg, m, n, k, monkeypatch):
    """The staged epilogue (TMEM -> stmatrix -> swizzled smem -> cp.async.bulk.tensor) moves the same BF16 values as the
    direct-store epilogue: ragged M / N edges are clipped by the tensor map, untouched memory stays untouched."""
    from deepgemm_b200 import _lib
    _, _, qa, qb = _quant_dense(m, n, k, seed=m + n)
    monkeypatch.setenv('DGB200_SPLITS', '1')
    outs = []
    for mode in ('0', '1'):
        monkeypatch.setenv('DGB200_TMA_STORE', mode)
        # D is a window of a larger buffer: anything written outside [m, n] shows up in the guard band
        buf = torch.full((m + 32, n + 64), 777.0, device='cuda', dtype=torch.bfloat16)
        d = buf[:m, :n]
        dg.fp8_gemm_nt(qa, qb, d)
        assert _lib.last_config()['tma_store'] == int(mode)
        assert bool((buf[m:] == 777.0).all()) and bool((buf[:, n:] == 777.0).all()), 'wrote outside D'
        outs.append(d.clone())
    assert torch.equal(outs[0], outs[1])


@pytest.mark.parametrize('block_m', [64, 128, 176, 240])
def test_tma_store_every_tile_height(dg, block_m, monkeypatch):
    m, n, k = 1000, 768, 640
    _, _, qa, qb = _quant_dense(m, n, k, seed=block_m)
    monkeypatch.setenv('DGB200_SPLITS', '1')
    monkeypatch.setenv('DGB200_TMA_STORE', '0')
    base = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    dg.fp8_gemm_nt(qa, qb, base)
    monkeypatch.setenv('DGB200_TMA_STORE', '1')
    monkeypatch.setenv('DGB200_BLOCK_M', str(block_m))
    d = torch.full_like(base, float('nan'))
    dg.fp8_gemm_nt(qa, qb, d)
    assert torch.equal(d, base)


def test_tma_store_contiguous_grouped(dg, monkeypatch):
    from deepgemm_b200 import _lib
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    gen = torch.Generator(device='cuda').manual_seed(21)
    g, n, k, per = 5, 640, 768, 128
    a = torch.randn((g * per, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qa = per_token_cast_to_fp8(a, True)
    b = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    q = [per_block_cast_to_fp8(b[i], True) for i in range(g)]
    qb = (torch.stack([x[0] for x in q]), torch.stack([x[1] for x in q]))
    layout = torch.arange(g, device='cuda', dtype=torch.int32).repeat_interleave(per)
    outs = []
    for mode in ('0', '1'):
        monkeypatch.setenv('DGB200_TMA_STORE', mode)
        d = torch.full((g * per, n), float('nan'), device='cuda', dtype=torch.bfloat16)
        dg.m_grouped_fp8_gemm_nt_contiguous(qa, qb, d, layout)
        assert _lib.last_config()['tma_store'] == int(mode)
        outs.append(d)
    assert torch.equal(outs[0], outs[1])


# ------------------------------------------------------------------------------------------------ pair split-K
@pytest.mark.parametrize('m,n,k,slices,bm', [(256, 4096, 7168, 2, 128), (384, 1024, 2048, 2, 192), (200, 768, 4096, 2, 224),
                                            (192, 2048, 7168, 4, 192), (256, 512, 2048, 4, 128), (130, 300, 1024, 2, 160)])
@pytest.mark.parametrize('out_dtype', [torch.bfloat16, torch.float32])
def test_pair_split_k_matches_oracle_and_is_deterministic(dg, m, n, k, slices, bm, out_dtype, monkeypatch):
    """K cut between the CTA pairs of a cluster (2 or 4 pairs), partial tiles reduce-scattered through distributed shared
    memory and added in slice order: FP32-rounding-level agreement with the one-pass kernel, run-to-run identical bits."""
    from deepgemm_b200 import _lib
    from oracle import blockwise
    from tests_helpers import assert_close_to_oracle
    _, _, qa, qb = _quant_dense(m, n, k, seed=m + k)
    monkeypatch.setenv('DGB200_CSPLIT', '0')
    monkeypatch.setenv('DGB200_PSPLIT', str(slices))
    monkeypatch.setenv('DGB200_PSPLIT_BM', str(bm))
    outs = []
    for _ in range(3):
        d = torch.full((m, n), float('nan'), device='cuda', dtype=out_dtype)
        dg.fp8_gemm_nt(qa, qb, d)
        cfg = _lib.last_config()
        assert cfg['cluster_split'] == slices and cfg['cluster'] == 2 * slices and cfg['block_m'] == bm, cfg
        outs.append(d)
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    assert_close_to_oracle(outs[0], blockwise.fp8_gemm_nt(_cpu(qa), _cpu(qb), out_dtype=out_dtype), f'psplit {slices} {m}x{n}x{k}')
    # accumulate into C
    c = (torch.randn((m, n), device='cuda') * 16).to(out_dtype)
    d = c.clone()
    dg.fp8_gemm_nt(qa, qb, d, c=d)
    want = blockwise.fp8_gemm_nt(_cpu(qa), _cpu(qb), out_dtype=out_dtype, c=c.cpu())
    prod = blockwise.fp8_gemm_nt(_cpu(qa), _cpu(qb), out_dtype=torch.float32)
    # BF16 accumulate = round(product) then one BF16 add: a sliced FP32 sum may round the product the other way AND move the
    # final rounding, i.e. up to two BF16 steps of (|product| + |C|) instead of one
    assert_close_to_oracle(d, want, 'psplit accumulate', mag=2 * (prod.abs() + c.cpu().float().abs()))


# ------------------------------------------------------------------------------------------------ second orientation
@pytest.mark.parametrize('m,n,k', [(64, 7168, 2048), (128, 24576, 1536), (100, 520, 768), (300, 1000, 512), (1, 2112, 7168),
                                   (256, 1001, 384), (64, 32768, 512)])
@pytest.mark.parametrize('out_dtype', [torch.bfloat16, torch.float32])
def test_transposed_output_orientation_gives_identical_bits(dg, m, n, k, out_dtype, monkeypatch):
    """Tokens on the TMEM lanes, weights tiled freely along N (fp8_gemm_kernel<..., kSwapD>): the same products summed over K
    in the same order, so the bits equal the weights-on-lanes kernel's; ragged N, unaligned row pitches (N = 1001) and C
    accumulation take the element-wise store path."""
    from deepgemm_b200 import _lib
    _, _, qa, qb = _quant_dense(m, n, k, seed=m * 7 + n)
    monkeypatch.setenv('DGB200_SPLITS', '1')
    outs = []
    for swap in ('0', '1'):
        monkeypatch.setenv('DGB200_SWAP', swap)
        buf = torch.full((m + 8, n + 24), 333.0, device='cuda', dtype=out_dtype)
        d = buf[:m, :n]
        dg.fp8_gemm_nt(qa, qb, d)
        cfg = _lib.last_config()
        assert cfg['swap_ab'] == int(swap)
        if swap == '1':
            assert cfg['cluster'] == (1 if m <= 128 else 2)
        assert bool((buf[m:] == 333.0).all()) and bool((buf[:, n:] == 333.0).all()), 'wrote outside D'
        outs.append(d.clone())
    assert torch.equal(outs[0], outs[1])
    c = (torch.randn((m, n), device='cuda') * 16).to(out_dtype)
    accs = []
    for swap in ('0', '1'):
        monkeypatch.setenv('DGB200_SWAP', swap)
        d = c.clone()
        dg.fp8_gemm_nt(qa, qb, d, c=d)
        accs.append(d)
    assert torch.equal(accs[0], accs[1])


@pytest.mark.parametrize('m,n,k,bn', [(512, 7168, 2048, 224), (300, 1000, 512, 96), (100, 520, 768, 32), (4096, 2048, 512, 224), (640, 1536, 256, 128),
                                      (257, 4104, 384, 160)])
def test_transposed_output_with_staged_tma_stores(dg, m, n, k, bn, monkeypatch):
    """Second orientation + per-warp staging through shared memory (32 rows x 32 columns turned around, written as 8 rows x 64
    contiguous bytes per instruction): same bits, nothing outside D."""
    from deepgemm_b200 import _lib
    _, _, qa, qb = _quant_dense(m, n, k, seed=m + bn)
    monkeypatch.setenv('DGB200_SPLITS', '1')
    monkeypatch.setenv('DGB200_SWAP', '0')
    base = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    dg.fp8_gemm_nt(qa, qb, base)
    monkeypatch.setenv('DGB200_SWAP', '1')
    monkeypatch.setenv('DGB200_TMA_STORE', '1')
    monkeypatch.setenv('DGB200_BLOCK_M', str(bn))
    buf = torch.full((m + 40, n + 72), 444.0, device='cuda', dtype=torch.bfloat16)
    d = buf[:m, :n]
    dg.fp8_gemm_nt(qa, qb, d)
    cfg = _lib.last_config()
    assert cfg['swap_ab'] == 1 and cfg['tma_store'] == 1 and cfg['block_m'] == bn
    assert bool((buf[m:] == 444.0).all()) and bool((buf[:, n:] == 444.0).all()), 'wrote outside D'
    assert torch.equal(d, base)


@pytest.mark.parametrize('bn', [16, 48, 112, 176, 240])
def test_transposed_output_every_tile_width(dg, bn, monkeypatch):
    m, n, k = 120, 2000, 640
    _, _, qa, qb = _quant_dense(m, n, k, seed=bn)
    monkeypatch.setenv('DGB200_SPLITS', '1')
    monkeypatch.setenv('DGB200_SWAP', '0')
    base = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    dg.fp8_gemm_nt(qa, qb, base)
    monkeypatch.setenv('DGB200_SWAP', '1')
    monkeypatch.setenv('DGB200_TMA_STORE', '0')
    monkeypatch.setenv('DGB200_BLOCK_M', str(bn))
    d = torch.full_like(base, float('nan'))
    dg.fp8_gemm_nt(qa, qb, d)
    assert torch.equal(d, base)


# ------------------------------------------------------------------------------------------------ skip_head_mid
@pytest.mark.parametrize('m,n,k,splits', [(128, 8192, 512, (128, 64, 128)), (4096, 2048, 512, (128, 64, 128)), (77, 768, 384, (64, 32, 128)),
                                          (33, 512, 256, (128, 0, 128))])
@pytest.mark.parametrize('out_dtype', [torch.bfloat16, torch.float32])
def test_skip_head_mid_places_every_column(dg, m, n, k, splits, out_dtype):
    """tests/test_attention.py:19-52 of the reference: D = concat over heads of [left | mid (untouched) | right]."""
    left, mid, right = splits
    _, _, qa, qb = _quant_dense(m, n, k, seed=n)
    heads = n // (left + right)
    plain = torch.empty((m, n), device='cuda', dtype=out_dtype)
    os.environ['DGB200_SPLITS'] = '1'
    try:
        dg.fp8_gemm_nt(qa, qb, plain)
        d = torch.full((m, n + heads * mid), 555.0, device='cuda', dtype=out_dtype)
        dg.fp8_gemm_nt_skip_head_mid(qa, qb, d, splits)
    finally:
        os.environ.pop('DGB200_SPLITS', None)
    dv, pv = d.view(m, heads, left + mid + right), plain.view(m, heads, left + right)
    assert torch.equal(dv[:, :, :left], pv[:, :, :left])
    assert torch.equal(dv[:, :, left + mid:], pv[:, :, left:])
    assert bool((dv[:, :, left:left + mid] == 555.0).all()), 'the gap must not be written'
    with pytest.raises(RuntimeError):
        dg.fp8_gemm_nt_skip_head_mid(qa, qb, plain, (left, mid + 1, right))        # attention.hpp:49


# ------------------------------------------------------------------------------------------------ einsum / bmm
def _per_batch_reference(dg, a, sfa, b, sfb, d_like, recipe, c=None):
    """The batched result, batch by batch, through the dense kernel (contiguous copies of every operand)."""
    outs = []
    for i in range(a.shape[0]):
        d = torch.empty(d_like.shape[1:], device='cuda', dtype=d_like.dtype)
        ci = None
        if c is not None:
            d.copy_(c[i])
            ci = d
        dg.fp8_gemm_nt((a[i], sfa[i]), (b[i], sfb[i]), d, c=ci, recipe=recipe)
        outs.append(d)
    return torch.stack(outs)


@pytest.mark.parametrize('b_', [4, 130])
def test_fp8_einsum_bhr_hdr_bhd(dg, b_):
    """tests/test_einsum.py:83-108 of the reference, (batch, m, n, k) = (h, b, d, r)."""
    from deepgemm_b200.testing import calc_diff
    from deepgemm_b200.utils import ceil_div, per_block_cast_to_fp8, per_token_cast_to_fp8
    h, r, dd = 8, 512, 384
    x = torch.randn((b_, h, r), device='cuda', dtype=torch.bfloat16)
    y = torch.randn((h, dd, r), device='cuda', dtype=torch.bfloat16)
    xq = per_token_cast_to_fp8(x.view(-1, r), True)
    xq = (xq[0].view(b_, h, r), xq[1].view(b_, h, ceil_div(r, 128)))
    yq = [per_block_cast_to_fp8(y[i], True) for i in range(h)]
    yq = (torch.stack([q[0] for q in yq]), torch.stack([q[1] for q in yq]))
    z = torch.full((b_, h, dd), float('nan'), device='cuda', dtype=torch.bfloat16)
    os.environ['DGB200_SPLITS'] = '1'
    try:
        dg.fp8_einsum('bhr,hdr->bhd', xq, yq, z)
        want = _per_batch_reference(dg, xq[0].permute(1, 0, 2), xq[1].permute(1, 0, 2), yq[0], yq[1], z.permute(1, 0, 2),
                                    (1, 128, 128))
    finally:
        os.environ.pop('DGB200_SPLITS', None)
    assert torch.equal(z.permute(1, 0, 2), want)
    assert calc_diff(z, torch.einsum('bhr,hdr->bhd', x, y)) < 1e-3


def test_fp8_einsum_bhd_hdr_bhr(dg):
    """tests/test_einsum.py:111-137: B operand MN-major through a permuted view."""
    from deepgemm_b200.testing import calc_diff
    from deepgemm_b200.utils import ceil_div, per_block_cast_to_fp8, per_token_cast_to_fp8
    b_, h, r, dd = 96, 4, 640, 256
    x = torch.randn((b_, h, dd), device='cuda', dtype=torch.bfloat16)
    y = torch.randn((h, dd, r), device='cuda', dtype=torch.bfloat16)
    xq = per_token_cast_to_fp8(x.view(-1, dd), True)
    xq = (xq[0].view(b_, h, dd), xq[1].view(b_, h, ceil_div(dd, 128)))
    yq = [per_block_cast_to_fp8(y[i], True) for i in range(h)]
    yq = (torch.stack([q[0] for q in yq]), torch.stack([q[1] for q in yq]))
    z = torch.full((b_, h, r), float('nan'), device='cuda', dtype=torch.bfloat16)
    dg.fp8_einsum('bhd,hdr->bhr', xq, yq, z)
    assert calc_diff(z, torch.einsum('bhd,hdr->bhr', x, y)) < 1e-3
    # bit-exact against the dense kernel on contiguous K-major copies of B[h]^T (same products, same K order)
    os.environ['DGB200_SPLITS'] = '1'
    try:
        for i in range(h):
            d = torch.empty((b_, r), device='cuda', dtype=torch.bfloat16)
            dg.fp8_gemm_nt((xq[0][:, i].contiguous(), xq[1][:, i].contiguous()),
                           (yq[0][i].t().contiguous(), yq[1][i].t().contiguous()), d)
            z2 = torch.empty((b_, h, r), device='cuda', dtype=torch.bfloat16)
            dg.fp8_einsum('bhd,hdr->bhr', xq, yq, z2)
            assert torch.equal(z2[:, i], d)
    finally:
        os.environ.pop('DGB200_SPLITS', None)


def test_fp8_einsum_bhd_bhr_hdr_accumulates_fp32(dg):
    """tests/test_einsum.py:140-165: both operands MN-major, K = the batch dim, FP32 accumulate into D, recipe (1, 1, 128)."""
    from deepgemm_b200.testing import calc_diff
    from deepgemm_b200.utils import ceil_div, per_channel_cast_to_fp8
    b_, h, r, dd = 512, 4, 256, 128
    x = torch.randn((b_, h, dd), device='cuda', dtype=torch.bfloat16)
    y = torch.randn((b_, h, r), device='cuda', dtype=torch.bfloat16)
    z0 = torch.randn((h, dd, r), device='cuda', dtype=torch.float32) * 10
    xq = per_channel_cast_to_fp8(x.view(b_, -1), True)
    yq = per_channel_cast_to_fp8(y.view(b_, -1), True)
    xq = (xq[0].view(b_, h, dd), xq[1].view(ceil_div(b_, 128), h, dd))
    yq = (yq[0].view(b_, h, r), yq[1].view(ceil_div(b_, 128), h, r))
    z = z0.clone()
    dg.fp8_einsum('bhd,bhr->hdr', xq, yq, z, z, recipe=(1, 1, 128))
    assert calc_diff(z, z0 + torch.einsum('bhd,bhr->hdr', x.float(), y.float())) < 1e-3
    # exact-input check: FP32 matmul of the dequantised operands
    xd = xq[0].float() * xq[1].repeat_interleave(128, 0)[:b_]
    yd = yq[0].float() * yq[1].repeat_interleave(128, 0)[:b_]
    torch.backends.cuda.matmul.allow_tf32 = False
    want = z0 + torch.einsum('bhd,bhr->hdr', xd, yd)
    assert ((z - want).abs().max() / want.abs().max()) < 1e-5
    with pytest.raises(RuntimeError):
        dg.fp8_einsum('bhd,bdr->hr', xq, yq, z)


# ------------------------------------------------------------------------------------------------ activation quantiser
@pytest.mark.parametrize('m,k', [(4096, 7168), (64, 7168), (1, 512), (130, 640), (37, 200), (5, 7296), (256, 96)])
@pytest.mark.parametrize('gran_k', [128, 32])
def test_cuda_quantiser_is_bit_identical_to_the_reference_python(dg, m, k, gran_k):
    """dgb200_per_token_cast_to_fp8 == per_token_cast_to_fp8(x, True, gran_k) (deep_gemm/utils/math.py:26-38; restated in
    deepgemm_b200/utils/math.py and pinned against the reference's own output by tests/test_oracle.py) followed by the
    MN-major packing of the GEMM's transform (csrc/apis/layout.hpp:48-58)."""
    from deepgemm_b200.utils import per_token_cast_to_fp8
    gen = torch.Generator(device='cuda').manual_seed(m * 31 + k)
    x = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen) * \
        torch.exp2(torch.randint(-12, 12, (m, 1), device='cuda', generator=gen).float()).to(torch.bfloat16)
    x[0, : min(k, 128)] = 0                                              # an all-zero block: amax clamps to 1e-4
    q, sf = dg.per_token_cast_to_fp8_packed(x, gran_k)
    q_ref, sf_ref = per_token_cast_to_fp8(x, True, gran_k)
    assert torch.equal(q.view(torch.uint8), q_ref.view(torch.uint8))
    want = dg.transform_sf_into_required_layout(sf_ref, m, k, (1, gran_k), None, None)
    assert sf.shape == want.shape and sf.stride() == want.stride()
    assert torch.equal(sf, want)
    if k % 16 == 0:
        # and the pair drops straight into the GEMM: identical output to the FP32-scale path
        from deepgemm_b200.utils import per_block_cast_to_fp8
        b = torch.randn((256, k), device='cuda', dtype=torch.bfloat16, generator=gen)
        qb = per_block_cast_to_fp8(b, True, 128)
        d0 = torch.empty((m, 256), device='cuda', dtype=torch.bfloat16)
        d1 = torch.empty_like(d0)
        dg.fp8_gemm_nt((q_ref, sf_ref), qb, d0, recipe_a=(1, gran_k), recipe_b=(128, 128))
        dg.fp8_gemm_nt((q, sf), qb, d1, recipe_a=(1, gran_k), recipe_b=(128, 128))
        assert torch.equal(d0, d1)


def test_cuda_quantiser_strided_input(dg):
    from deepgemm_b200.utils import per_token_cast_to_fp8
    big = torch.randn((64, 1024 + 24), device='cuda', dtype=torch.bfloat16)
    for x in (big[:, :1024], big[:, 3:1003]):                            # aligned pitch / unaligned base
        q, sf = dg.per_token_cast_to_fp8_packed(x)
        q_ref, sf_ref = per_token_cast_to_fp8(x.contiguous(), True)
        assert torch.equal(q.view(torch.uint8), q_ref.view(torch.uint8))
        assert torch.equal(sf, dg.transform_sf_into_required_layout(sf_ref, 64, x.shape[1], (1, 128), None, None))


# ------------------------------------------------------------------------------------------------ gran_k 32 / mixed recipes
@pytest.mark.parametrize('recipe_a,recipe_b', [((1, 32), (128, 32)), ((1, 32), (128, 128)), ((1, 128), (1, 32)), ((1, 32), (1, 32))])
@pytest.mark.parametrize('m,n,k', [(200, 384, 1024), (64, 4096, 7168)])
def test_dense_gran_k_32_and_mixed_recipes_match_oracle(dg, recipe_a, recipe_b, m, n, k):
    """csrc/apis/layout.hpp:24-35,74-88: `recipe_a` / `recipe_b` = (gran_mn, gran_k) per operand; gran_k in {32, 128} on SM100."""
    from oracle import blockwise
    from tests_helpers import assert_close_to_oracle
    gen = torch.Generator(device='cuda').manual_seed(m + k + recipe_a[1] + recipe_b[1])
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qa, qb = _cast(a, recipe_a), _cast(b, recipe_b)
    d = torch.full((m, n), float('nan'), device='cuda', dtype=torch.bfloat16)
    dg.fp8_gemm_nt(qa, qb, d, recipe_a=recipe_a, recipe_b=recipe_b)
    ad = blockwise.dequant(qa[0].cpu(), qa[1].cpu(), recipe_a[0], recipe_a[1])
    bd = blockwise.dequant(qb[0].cpu(), qb[1].cpu(), recipe_b[0], recipe_b[1])
    want = (ad.double() @ bd.double().t()).float().to(torch.bfloat16)
    assert_close_to_oracle(d, want, f'{recipe_a} {recipe_b}')


def _cast(x, recipe):
    """FP8 + FP32 power-of-two scales at (gran_mn, gran_k) granularity."""
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    gran_mn, gran_k = recipe
    if gran_mn == 1:
        return per_token_cast_to_fp8(x, True, gran_k)
    assert gran_mn == 128
    if gran_k == 128:
        return per_block_cast_to_fp8(x, True, 128)
    # 128 x 32 blocks: amax over 128 rows x 32 columns
    from deepgemm_b200.utils.math import _scale_from_amax
    rows, cols = x.shape
    rp, cp = -(-rows // 128) * 128, -(-cols // 32) * 32
    xp = x.new_zeros((rp, cp))
    xp[:rows, :cols] = x
    blocks = xp.view(rp // 128, 128, cp // 32, 32)
    sf = _scale_from_amax(blocks.abs().float().amax(dim=(1, 3), keepdim=True), True)
    q = (blocks * (1.0 / sf)).to(torch.float8_e4m3fn).view(rp, cp)[:rows, :cols].contiguous()
    return q, sf.view(rp // 128, cp // 32)


def test_m_grouped_gran_k_32_matches_oracle(dg):
    from oracle import blockwise
    from tests_helpers import assert_close_to_oracle
    from deepgemm_b200.utils import per_token_cast_to_fp8
    gen = torch.Generator(device='cuda').manual_seed(17)
    g, n, k, per = 3, 384, 512, 128
    a = torch.randn((g * per, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qa = per_token_cast_to_fp8(a, True, 32)
    q = [per_token_cast_to_fp8(b[i], True, 32) for i in range(g)]
    qb = (torch.stack([x[0] for x in q]), torch.stack([x[1] for x in q]))
    layout = torch.arange(g, device='cuda', dtype=torch.int32).repeat_interleave(per)
    d = torch.full((g * per, n), float('nan'), device='cuda', dtype=torch.bfloat16)
    dg.m_grouped_fp8_gemm_nt_contiguous(qa, qb, d, layout, recipe=(1, 1, 32))
    want, valid = blockwise.m_grouped_fp8_gemm_nt_contiguous(_cpu(qa), _cpu(qb), layout.cpu(), recipe=(1, 1, 32))
    assert_close_to_oracle(d.cpu()[valid], want[valid], 'contiguous gran_k 32')


# ------------------------------------------------------------------------------------------------ full-size digests
def _run_normal(dg, case):
    import cases
    from deepgemm_b200 import utils
    qa, qb, c, d = cases.make_normal(case, utils)
    if c is not None:
        d.copy_(c)
    dg.fp8_gemm_nt(qa, qb, d, c=d if c is not None else None)
    torch.cuda.synchronize()
    return qa, qb, c, d


def _normal_case_ids():
    import cases
    return [c['name'] for c in cases.normal_cases()]


@pytest.mark.parametrize('name', _normal_case_ids())
def test_reference_shape_list_matches_the_reference_kernel_bit_for_bit(dg, digests, name, monkeypatch):
    """The reference's own `enumerate_normal` list (tests/generators.py:115-154): forward M in {1, 128, 4096} x its
    (N, K) pairs, BF16 accumulation, and the MN-major dgrad / wgrad forms (FP32 accumulate). With split-K off our
    output bytes hash to the digest of the reference kernel's output on the same inputs."""
    import cases
    case = next(c for c in cases.normal_cases() if c['name'] == name)
    if name not in digests:
        pytest.skip(f'no digest for {name}')
    monkeypatch.setenv('DGB200_SPLITS', '1')
    _, _, _, d = _run_normal(dg, case)
    assert cases.digest(d) == digests[name], name


@pytest.mark.parametrize('name', ['fwd_1x2112x7168', 'fwd_128x576x7168', 'fwd_128x7168x2048'])
def test_reference_shape_list_default_split_k_is_within_tolerance(dg, name):
    """Default configuration (cluster split-K on for M <= 128): FP32-rounding-level agreement with the exact-input FP32
    matmul, the stated tolerance of DESIGN.md section 2."""
    import cases
    case = next(c for c in cases.normal_cases() if c['name'] == name)
    qa, qb, _, d = _run_normal(dg, case)
    torch.backends.cuda.matmul.allow_tf32 = False
    ref = (qa[0].float() * qa[1].repeat_interleave(128, 1)[:, :case['k']]) @ \
          (qb[0].float() * qb[1].repeat_interleave(128, 0)[:case['n']].repeat_interleave(128, 1)[:, :case['k']]).t()
    from deepgemm_b200.testing import calc_diff
    assert calc_diff(d, ref.to(torch.bfloat16)) < 1e-6          # (the BF16 rounding itself is ~1.3e-6 against the FP32 value)
    err = (d.float() - ref).abs()
    assert bool((err <= ref.abs() * 2.0 ** -7 + 1e-5 * ref.abs().max()).all())


@pytest.mark.parametrize('mean_m', [64, 128])
def test_contiguous_g256_full_size_matches_the_reference_kernel(dg, digests, mean_m):
    """BASELINE config 3: 256 experts, N=4096, K=7168, variable M. Digest of the valid rows == the reference kernel's;
    a sample of experts is also checked against the FP64-accumulated oracle."""
    import cases
    from deepgemm_b200 import utils
    from oracle import blockwise
    from tests_helpers import assert_close_to_oracle
    p = cases.make_contiguous(mean_m, utils)
    d = torch.zeros((p['m'], 4096), device='cuda', dtype=torch.bfloat16)
    dg.m_grouped_fp8_gemm_nt_contiguous(p['a'], p['b'], d, p['layout'])
    torch.cuda.synchronize()
    key = f'contiguous_g256_m{mean_m}'
    if key in digests:
        assert cases.digest(d[p['valid']]) == digests[key]
    layout = p['layout'].cpu()
    for e in (0, 101, 255):
        rows = (layout == e).nonzero().flatten()[:48]
        ad = blockwise.dequant(p['a'][0][rows.cuda()].cpu(), p['a'][1][rows.cuda()].cpu(), 1, 128)
        bd = blockwise.dequant(p['b'][0][e].cpu(), p['b'][1][e].cpu(), 128, 128)
        want = (ad.double() @ bd.double().t()).float().to(torch.bfloat16)
        assert_close_to_oracle(d[rows.cuda()], want, f'contiguous expert {e}')
    if key not in digests:
        pytest.skip('oracle sample passed; no reference digest yet')


@pytest.mark.parametrize('mean_m', [16, 64, 96])
def test_masked_g256_full_size_matches_the_reference_kernel(dg, digests, mean_m):
    """BASELINE config 4: 256 experts, M_max=128, N=7168, K=2048, under a CUDA graph like the decode path."""
    import cases
    from deepgemm_b200 import utils
    from oracle import blockwise
    from tests_helpers import assert_close_to_oracle
    p = cases.make_masked(mean_m, utils)
    sfa = dg.transform_sf_into_required_layout(p['a'][1], 128, 2048, (1, 128, 128), 256, True)
    sfb = dg.transform_sf_into_required_layout(p['b'][1], 7168, 2048, (1, 128, 128), 256, False)
    d = torch.zeros((256, 128, 7168), device='cuda', dtype=torch.bfloat16)
    call = lambda: dg.m_grouped_fp8_gemm_nt_masked((p['a'][0], sfa), (p['b'][0], sfb), d, p['masked_m'], p['expected_m'])  # noqa: E731
    call()
    graph, side = torch.cuda.CUDAGraph(), torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            call()
    d.zero_()
    graph.replay()
    torch.cuda.synchronize()
    key = f'masked_g256_m{mean_m}'
    if key in digests:
        assert cases.digest_masked(d, p['masked_m']) == digests[key]
    masked = p['masked_m'].tolist()
    for e in (0, 77, 255):
        mg = masked[e]
        if mg == 0:
            continue
        ad = blockwise.dequant(p['a'][0][e, :mg].cpu(), p['a'][1][e, :mg].cpu(), 1, 128)
        bd = blockwise.dequant(p['b'][0][e, :1024].cpu(), p['b'][1][e, :8].cpu(), 128, 128)
        want = (ad.double() @ bd.double().t()).float().to(torch.bfloat16)
        assert_close_to_oracle(d[e, :mg, :1024], want, f'masked expert {e}')
        assert bool((d[e, mg:] == 0).all())
    if key not in digests:
        pytest.skip('oracle sample passed; no reference digest yet')


# ------------------------------------------------------------------------------------------------ psum robustness
def test_psum_layout_with_a_buffer_that_is_not_a_multiple_of_the_alignment(dg):
    """A psum walk whose aligned group start runs past `m` (an EP buffer sized in 16-row units, or prefix sums clamped on
    overflow) must stop at the buffer end: short segments, no rows >= m, no endless tile loop."""
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    from oracle import blockwise
    from tests_helpers import assert_close_to_oracle
    gen = torch.Generator(device='cuda').manual_seed(3)
    g, n, k, m = 4, 256, 512, 1008                      # 1008 = 7 * 128 + 112
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qa = per_token_cast_to_fp8(a, True)
    b = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    q = [per_block_cast_to_fp8(b[i], True) for i in range(g)]
    qb = (torch.stack([x[0] for x in q]), torch.stack([x[1] for x in q]))
    # expert 0: rows [0, 1000); expert 1 would start at 1024 > m; experts 2, 3 clamp to m as an overflowing dispatch does
    psum = torch.tensor([1000, 1008, 1008, 1008], device='cuda', dtype=torch.int32)
    buf = torch.full((m + 256, n), 321.0, device='cuda', dtype=torch.bfloat16)
    d = buf[:m]
    dg.m_grouped_fp8_gemm_nt_contiguous(qa, qb, d, psum, use_psum_layout=True)
    torch.cuda.synchronize()
    assert bool((buf[m:] == 321.0).all()), 'wrote past the end of D'
    ad = blockwise.dequant(qa[0][:1000].cpu(), qa[1][:1000].cpu(), 1, 128)
    bd = blockwise.dequant(qb[0][0].cpu(), qb[1][0].cpu(), 128, 128)
    assert_close_to_oracle(d[:1000], (ad.double() @ bd.double().t()).float().to(torch.bfloat16), 'expert 0')
    assert bool((d[1000:] == 0).all())                  # zero padding up to the (clamped) aligned end
