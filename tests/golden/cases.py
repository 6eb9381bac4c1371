"""Seeded full-size problem generators shared by the golden-digest script (tests/golden/make_golden_digests.py, which runs
the reference's own SM100 kernel on them) and the GPU parity tests (which run ours and compare SHA-256 digests of the
output bytes). Inputs are generated on the device from a seeded generator with this repository's quantisers (bit-identical
to the reference's, tests/test_oracle.py), so the same bytes come out on every B200 with this image's torch.

Shapes follow the reference's own test generators (tests/generators.py:115-154 `enumerate_normal`, :332-355 contiguous,
:387-406 masked) and BASELINE.json configs 3 / 4.
"""
import hashlib
import random

import torch

# (m, n, k, major_a, major_b, accumulate, out dtype) -- 'k' = K-major, 'mn' = MN-major
NK_FWD = [(2112, 7168), (576, 7168), (24576, 1536), (32768, 512), (7168, 16384), (7168, 2048)]
NK_BWD = [(2112, 7168), (7168, 2048)]


def normal_cases():
    cases = []
    for m in (1, 128, 4096):
        for n, k in NK_FWD:
            cases.append(dict(name=f'fwd_{m}x{n}x{k}', m=m, n=n, k=k, major_a='k', major_b='k', acc=False, dtype='bf16'))
    cases.append(dict(name='fwd_acc_128x2112x7168', m=128, n=2112, k=7168, major_a='k', major_b='k', acc=True, dtype='bf16'))
    cases.append(dict(name='fwd_acc_4096x7168x2048', m=4096, n=7168, k=2048, major_a='k', major_b='k', acc=True, dtype='bf16'))
    for n, k in NK_BWD:                                   # tests/generators.py:143-154 with m = 4096
        m = 4096
        cases.append(dict(name=f'dgrad_{m}x{k}x{n}', m=m, n=k, k=n, major_a='k', major_b='mn', acc=False, dtype='bf16'))
        cases.append(dict(name=f'wgrad_f32_{n}x{m}x{k}', m=n, n=m, k=k, major_a='mn', major_b='mn', acc=True, dtype='f32'))
        cases.append(dict(name=f'wgrad_bf16_{n}x{m}x{k}', m=n, n=m, k=k, major_a='mn', major_b='mn', acc=False, dtype='bf16'))
    return cases


def _seed_of(name: str) -> int:
    return int.from_bytes(hashlib.sha256(name.encode()).digest()[:4], 'little')


def make_normal(case, utils):
    """-> (a pair, b pair, c or None, d). MN-major operands are produced like the reference's generator does
    (generators.py:49-52, 248-254): quantise the K-major tensor, then store it transposed."""
    gen = torch.Generator(device='cuda').manual_seed(_seed_of(case['name']))
    m, n, k = case['m'], case['n'], case['k']
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qa, qb = utils.per_token_cast_to_fp8(a, True), utils.per_block_cast_to_fp8(b, True)
    del a, b
    if case['major_a'] == 'mn':
        qa = (qa[0].t().contiguous().t(), qa[1])
    if case['major_b'] == 'mn':
        qb = (qb[0].t().contiguous().t(), qb[1])
    dtype = torch.bfloat16 if case['dtype'] == 'bf16' else torch.float32
    c = (torch.randn((m, n), device='cuda', dtype=torch.float32, generator=gen) * 32).to(dtype) if case['acc'] else None
    d = torch.empty((m, n), device='cuda', dtype=dtype)
    return qa, qb, c, d


def grouped_weights(g, n, k, gen, utils):
    b = torch.empty((g, n, k), device='cuda', dtype=torch.float8_e4m3fn)
    sfb = torch.empty((g, (n + 127) // 128, (k + 127) // 128), device='cuda', dtype=torch.float32)
    for i in range(g):
        b[i], sfb[i] = utils.per_block_cast_to_fp8(torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen), True)
    return b, sfb


def make_contiguous(mean_m, utils, g=256, n=4096, k=7168, alignment=128, weights=None):
    """BASELINE config 3. -> dict(a pair, b pair, layout [M] int32, valid rows mask, m)"""
    rnd = random.Random(1000 + mean_m)
    gen = torch.Generator(device='cuda').manual_seed(3000 + mean_m)
    ms = [int(mean_m * rnd.uniform(0.7, 1.3)) for _ in range(g)]
    aligned = [(x + alignment - 1) // alignment * alignment for x in ms]
    m = sum(aligned)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    layout = torch.empty(m, dtype=torch.int32)
    s = 0
    for i, (mi, ai) in enumerate(zip(ms, aligned)):
        layout[s:s + mi] = i
        layout[s + mi:s + ai] = -1
        s += ai
    layout = layout.cuda()
    a[layout < 0] = 0
    qa = utils.per_token_cast_to_fp8(a, True)
    del a
    if weights is None:
        weights = grouped_weights(g, n, k, torch.Generator(device='cuda').manual_seed(77), utils)
    return dict(a=qa, b=weights, layout=layout, valid=layout >= 0, m=m, ms=ms, alignment=alignment)


def make_masked(mean_m, utils, g=256, m_max=128, n=7168, k=2048, weights=None):
    """BASELINE config 4. -> dict(a pair [G,M_max,K], b pair, masked_m, expected_m)"""
    rnd = random.Random(2000 + mean_m)
    gen = torch.Generator(device='cuda').manual_seed(4000 + mean_m)
    a = torch.randn((g, m_max, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    q = utils.per_token_cast_to_fp8(a.view(g * m_max, k), True)
    qa = (q[0].view(g, m_max, k), q[1].view(g, m_max, -1))
    masked = torch.tensor([min(m_max, int(mean_m * rnd.uniform(0.7, 1.3))) for _ in range(g)], dtype=torch.int32).cuda()
    if weights is None:
        weights = grouped_weights(g, n, k, torch.Generator(device='cuda').manual_seed(78), utils)
    return dict(a=qa, b=weights, masked_m=masked, expected_m=int(1.2 * mean_m))


BF16_CASES = [(128, 2112, 7168), (4096, 7168, 2048), (64, 576, 7168)]


def make_bf16(m, n, k):
    gen = torch.Generator(device='cuda').manual_seed(_seed_of(f'bf16_{m}x{n}x{k}'))
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    return a, b


def digest(t: torch.Tensor) -> str:
    """SHA-256 of the tensor's bytes in row-major order."""
    return hashlib.sha256(t.contiguous().cpu().view(torch.uint8).numpy().tobytes()).hexdigest()


def digest_masked(d: torch.Tensor, masked_m: torch.Tensor) -> str:
    rows = torch.arange(d.shape[1], device=d.device).unsqueeze(0) < masked_m.to(d.device).unsqueeze(1)   # [G, M_max]
    return digest(d[rows])
