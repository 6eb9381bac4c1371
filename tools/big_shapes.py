"""Largest-size sanity against the reference's kernel (32-bit index arithmetic, tensor-map extents): a few very tall / very wide dense
problems and one large contiguous grouped call, bit-compared on the same tensors. Development / evidence tool."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference  # noqa: E402
import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8  # noqa: E402

ref = import_reference()
dg.set_split_k(False)
gen = torch.Generator(device='cuda').manual_seed(3)
for (m, n, k) in [(131072, 4096, 512), (262144, 2048, 256), (16, 262144, 512), (70000, 7168, 384)]:
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
    qa, qb = per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True)
    del a, b
    sfa = ref.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
    sfb = ref.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
    d0 = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    d1 = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    ref.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d0)
    dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d1)
    torch.cuda.synchronize()
    print(json.dumps({'dense': [m, n, k], 'output_bytes': m * n * 2, 'bitwise_equal': bool(torch.equal(d0, d1))}), flush=True)
    del d0, d1, qa, qb, sfa, sfb
    torch.cuda.empty_cache()
# contiguous: 64 experts, ~2300 rows each (sum M ~ 150k), N = 4096, K = 1024
g, n, k, align = 64, 4096, 1024, dg.get_mk_alignment_for_contiguous_layout()
ws = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
qs = [per_block_cast_to_fp8(ws[i], True) for i in range(g)]
wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
del ws, qs
sfw = ref.transform_sf_into_required_layout(wq[1], n, k, (1, 128, 128), g, False)
ms = [2300 + 37 * (i % 7) for i in range(g)]
al = [(x + align - 1) // align * align for x in ms]
mt = sum(al)
at = torch.randn((mt, k), device='cuda', dtype=torch.bfloat16, generator=gen)
qat = per_token_cast_to_fp8(at, True)
del at
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
torch.cuda.synchronize()
print(json.dumps({'contiguous': {'groups': g, 'sum_m': mt, 'n': n, 'k': k}, 'bitwise_equal_valid_rows': bool(torch.equal(g0[valid], g1[valid]))}), flush=True)
