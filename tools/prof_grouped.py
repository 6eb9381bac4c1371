"""Run one contiguous grouped GEMM a few times (for ncu). usage: prof_grouped.py {ours|ref} G mean_m [iters]"""
import os
import random
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference, make_grouped_weights  # noqa: E402

which, g, mean_m = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 3
n, k, alignment = 4096, 7168, 128
random.seed(0)
lib = import_reference() if which == 'ref' else __import__('deepgemm_b200')
from deepgemm_b200.utils import per_token_cast_to_fp8  # noqa: E402
b, sfb = make_grouped_weights(g, n, k)
sfb_p = lib.transform_sf_into_required_layout(sfb, n, k, (1, 128, 128), g, False)
ms = [int(mean_m * random.uniform(0.7, 1.3)) for _ in range(g)]
aligned = [(x + alignment - 1) // alignment * alignment for x in ms]
m = sum(aligned)
a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
layout = torch.empty(m, device='cuda', dtype=torch.int32)
s0 = 0
for i, (mi, ai) in enumerate(zip(ms, aligned)):
    layout[s0:s0 + mi] = i
    layout[s0 + mi:s0 + ai] = -1
    a[s0 + mi:s0 + ai] = 0
    s0 += ai
qa = per_token_cast_to_fp8(a, True)
sfa = lib.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
flush = torch.empty(256 << 20, dtype=torch.int32, device='cuda')
for _ in range(iters):
    flush.zero_()
    lib.m_grouped_fp8_gemm_nt_contiguous((qa[0], sfa), (b, sfb_p), d, layout)
torch.cuda.synchronize()
print('done', which, g, mean_m, m)
