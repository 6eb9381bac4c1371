"""Randomised bit-for-bit comparison with the unmodified reference kernels (oracle/_ref) on the same tensors: dense FP8 (random M,
all major combinations, BF16 / FP32 / accumulate), grouped contiguous and masked FP8 with random segment sizes, BF16 dense.
set_split_k(False): the one documented deviation (K slices added in slice order) is switched off, everything else runs the
library's default heuristics. Time-boxed: the reference compiles one kernel per tile configuration. Development / evidence tool."""
import json
import os
import random
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference  # noqa: E402
import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8  # noqa: E402

BUDGET = float(os.environ.get('FUZZ_SECONDS', '240'))
ref = import_reference()
rng = random.Random(2024)
gen = torch.Generator(device='cuda').manual_seed(11)
t0 = time.time()
stats = {'dense_fp8': [0, 0], 'dense_fp8_majors': [0, 0], 'dense_fp8_fp32_accumulate': [0, 0], 'contiguous_fp8': [0, 0], 'masked_fp8': [0, 0], 'dense_bf16': [0, 0]}
failures = []
dg.set_split_k(False)


def note(kind, ok, what):
    stats[kind][0] += 1
    stats[kind][1] += int(ok)
    if not ok:
        failures.append((kind, what))


def rand_m():
    r = rng.random()
    if r < 0.4:
        return rng.randint(1, 160)
    if r < 0.8:
        return rng.randint(161, 1200)
    return rng.choice([2048, 3000, 4096, 4100])


NK = [(4096, 7168), (2112, 7168), (7168, 2048), (576, 7168), (1536, 1536), (24576, 1536)]
weights = {}
for (n, k) in NK:
    w = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qb = per_block_cast_to_fp8(w, True)
    weights[(n, k)] = (w, qb, ref.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False))

rounds = 0
while time.time() - t0 < BUDGET:
    rounds += 1
    n, k = NK[rounds % len(NK)]
    w, qb, sfb = weights[(n, k)]
    m = rand_m()
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qa = per_token_cast_to_fp8(a, True)
    sfa = ref.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
    d0, d1 = torch.empty((m, n), device='cuda', dtype=torch.bfloat16), torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    ref.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d0)
    dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d1)
    note('dense_fp8', torch.equal(d0, d1), (m, n, k))
    if rounds % 3 == 0 and n <= 7168:
        c = torch.randn((m, n), device='cuda', dtype=torch.float32, generator=gen)
        if m >= 16:
            # MN-major operands + FP32 accumulation (the dgrad / wgrad forms); an MN-major A needs 16-byte rows: M % 16 == 0
            mm = m // 16 * 16
            a_mn, b_mn = qa[0][:mm].t().contiguous().t(), qb[0].t().contiguous().t()
            sfa_mm = ref.transform_sf_into_required_layout(qa[1][:mm].contiguous(), mm, k, (1, 128, 128), None, True)
            e0, e1 = c[:mm].clone(), c[:mm].clone()
            ref.fp8_gemm_nt((a_mn, sfa_mm), (b_mn, sfb), e0, c=e0)
            dg.fp8_gemm_nt((a_mn, sfa_mm), (b_mn, sfb), e1, c=e1)
            note('dense_fp8_majors', torch.equal(e0, e1), (mm, n, k, 'mn-major a, b + fp32 c'))
        e0, e1 = c.clone(), c.clone()
        ref.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), e0, c=e0)
        dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), e1, c=e1)
        note('dense_fp8_fp32_accumulate', torch.equal(e0, e1), (m, n, k))
    if rounds % 4 == 0 and n <= 7168:
        ab = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
        f0, f1 = torch.empty((m, n), device='cuda', dtype=torch.bfloat16), torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
        ref.bf16_gemm_nt(ab, w, f0)
        dg.bf16_gemm_nt(ab, w, f1)
        note('dense_bf16', torch.equal(f0, f1), (m, n, k))
    if rounds % 5 == 0 and n <= 4096:
        g = rng.randint(2, 6)
        ws = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
        qs = [per_block_cast_to_fp8(ws[i], True) for i in range(g)]
        wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
        sfw = ref.transform_sf_into_required_layout(wq[1], n, k, (1, 128, 128), g, False)
        align = dg.get_mk_alignment_for_contiguous_layout()
        ms = [rng.randint(0, 300) for _ in range(g)]
        al = [(x + align - 1) // align * align for x in ms]
        mt = max(sum(al), align)
        at = torch.randn((mt, k), device='cuda', dtype=torch.bfloat16, generator=gen)
        qat = per_token_cast_to_fp8(at, True)
        sfat = ref.transform_sf_into_required_layout(qat[1], mt, k, (1, 128, 128), None, True)
        layout = torch.full((mt,), -1, device='cuda', dtype=torch.int32)
        s = 0
        for i, (mi, ai) in enumerate(zip(ms, al)):
            layout[s:s + mi] = i
            s += ai
        valid = layout >= 0
        g0, g1 = torch.zeros((mt, n), device='cuda', dtype=torch.bfloat16), torch.zeros((mt, n), device='cuda', dtype=torch.bfloat16)
        ref.m_grouped_fp8_gemm_nt_contiguous((qat[0], sfat), (wq[0], sfw), g0, layout)
        dg.m_grouped_fp8_gemm_nt_contiguous((qat[0], sfat), (wq[0], sfw), g1, layout)
        note('contiguous_fp8', torch.equal(g0[valid], g1[valid]), (ms, n, k))
        mmax = 128
        am = torch.randn((g, mmax, k), device='cuda', dtype=torch.bfloat16, generator=gen)
        qam = [per_token_cast_to_fp8(am[i], True) for i in range(g)]
        qam = (torch.stack([q[0] for q in qam]), torch.stack([q[1] for q in qam]))
        sfam = ref.transform_sf_into_required_layout(qam[1], mmax, k, (1, 128, 128), g, True)
        masked = torch.tensor([rng.randint(0, mmax) for _ in range(g)], device='cuda', dtype=torch.int32)
        h0, h1 = torch.zeros((g, mmax, n), device='cuda', dtype=torch.bfloat16), torch.zeros((g, mmax, n), device='cuda', dtype=torch.bfloat16)
        exp_m = max(1, int(masked.float().mean()))
        ref.m_grouped_fp8_gemm_nt_masked((qam[0], sfam), (wq[0], sfw), h0, masked, exp_m)
        dg.m_grouped_fp8_gemm_nt_masked((qam[0], sfam), (wq[0], sfw), h1, masked, exp_m)
        ok = all(torch.equal(h0[i, :int(masked[i])], h1[i, :int(masked[i])]) for i in range(g))
        note('masked_fp8', ok, (masked.tolist(), n, k))
    torch.cuda.synchronize()
dg.set_split_k(True)
print(json.dumps({'seconds': round(time.time() - t0, 1), 'rounds': rounds, 'cases_run_and_bit_identical': {k: v for k, v in stats.items()},
                  'failures': failures[:20]}), flush=True)
sys.exit(1 if failures else 0)
