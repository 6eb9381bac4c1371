#!/usr/bin/env bash
mkdir -p gpurun_out
cat > /tmp/mc.py <<'PY'
import torch, sys
sys.path.insert(0, '.')
import deepgemm_b200 as dg
from deepgemm_b200 import ep, _lib
from deepgemm_b200.utils import per_token_cast_to_fp8, per_block_cast_to_fp8
import os
def dense(m, n, k, env=None, c=False, fp32=False):
    for kk in ('DGB200_CSPLIT', 'DGB200_SPLITS', 'DGB200_BLOCK_M'): os.environ.pop(kk, None)
    if env: os.environ.update(env)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16); b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16)
    qa, qb = per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True)
    d = torch.zeros((m, n), device='cuda', dtype=torch.float32 if fp32 else torch.bfloat16)
    dg.fp8_gemm_nt(qa, qb, d, c=d if c else None)
    torch.cuda.synchronize()
    print('dense', m, n, k, env, _lib.last_config())
dense(100, 520, 1536, {'DGB200_CSPLIT': '4'})
dense(33, 136, 1408, {'DGB200_CSPLIT': '2'}, c=True)
dense(1, 2112, 7168)
dense(300, 2112, 1536, fp32=True)
dense(1100, 1000, 640)            # balanced heights, ragged N
dense(64, 768, 2048, {'DGB200_SPLITS': '4'})
# grouped psum
g, n, k = 4, 256, 512
w = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16)
qs = [per_block_cast_to_fp8(w[i], True) for i in range(g)]
wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
x = torch.randn((333, k), device='cuda', dtype=torch.bfloat16)
xq, sfp = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
ids = torch.randint(0, g, (333,), device='cuda'); ids[::9] = -1
buf = ep.EpBuffer(g, 333 + g * 128, k)
d = buf.output(n)
for ov in (False, True):
    _, r = ep.expert_sharded_grouped_gemm(xq, sfp, ids, wq, buf, d, overlap=ov)
    out = buf.combine(r.token_row, ids)
    torch.cuda.synchronize()
buf.close()
# masked
a = torch.randn((g, 64, k), device='cuda', dtype=torch.bfloat16)
qa = [per_token_cast_to_fp8(a[i], True) for i in range(g)]
qa = (torch.stack([q[0] for q in qa]), torch.stack([q[1] for q in qa]))
dm = torch.zeros((g, 64, n), device='cuda', dtype=torch.bfloat16)
dg.m_grouped_fp8_gemm_nt_masked(qa, wq, dm, torch.tensor([5, 64, 0, 33], device='cuda', dtype=torch.int32), 32)
torch.cuda.synchronize()
print('memcheck script done')
PY
timeout 900 compute-sanitizer --tool memcheck --report-api-errors no --error-exitcode 7 python /tmp/mc.py > gpurun_out/memcheck.log 2>&1; echo "memcheck rc=$?"; grep -c "Invalid\|ERROR SUMMARY" gpurun_out/memcheck.log; tail -5 gpurun_out/memcheck.log | cut -c1-300; grep -m5 -A12 "Invalid" gpurun_out/memcheck.log | cut -c1-200
