"""One compact pass over every kernel variant of the library, for compute-sanitizer (tools/gpu_sanitize.sh runs it under
memcheck, synccheck and racecheck like the reference's tests/test_sanitizer.py:52-79 does for its kernels). Small shapes:
the sanitizers slow kernels down 10-100x."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200 import _lib, ep  # noqa: E402
from deepgemm_b200.utils import per_block_cast_to_fp8, per_channel_cast_to_fp8, per_token_cast_to_fp8  # noqa: E402

KNOBS = ('DGB200_CSPLIT', 'DGB200_SPLITS', 'DGB200_BLOCK_M', 'DGB200_PSPLIT', 'DGB200_PSPLIT_BM', 'DGB200_TMA_STORE', 'DGB200_CLUSTER')


def dense(m, n, k, env=None, c=False, fp32=False, majors='kk'):
    for kk in KNOBS:
        os.environ.pop(kk, None)
    if env:
        os.environ.update(env)
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16)
    qa, qb = per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True)
    if majors[0] == 'm':
        qa = (qa[0].t().contiguous().t(), qa[1])
    if majors[1] == 'm':
        qb = (qb[0].t().contiguous().t(), qb[1])
    d = torch.zeros((m, n), device='cuda', dtype=torch.float32 if fp32 else torch.bfloat16)
    dg.fp8_gemm_nt(qa, qb, d, c=d if c else None)
    torch.cuda.synchronize()
    print('dense', m, n, k, env, majors, _lib.last_config(), flush=True)


dense(100, 520, 1536, {'DGB200_CSPLIT': '4'})
dense(33, 136, 1408, {'DGB200_CSPLIT': '2'}, c=True)
if 'LONG_K' in os.environ:
    dense(1, 2112, 7168)
dense(300, 2112, 1536, fp32=True)
dense(1100, 1000, 640, {'DGB200_TMA_STORE': '1'})                       # staged TMA-store epilogue, ragged M and N
dense(520, 512, 768, {'DGB200_TMA_STORE': '1', 'DGB200_BLOCK_M': '240'})
dense(64, 768, 2048, {'DGB200_SPLITS': '4'})                            # workspace split-K
dense(256, 512, 2048, {'DGB200_CSPLIT': '0', 'DGB200_PSPLIT': '2', 'DGB200_PSPLIT_BM': '128'})     # pair split-K, cluster of 4
dense(192, 512, 2048, {'DGB200_CSPLIT': '0', 'DGB200_PSPLIT': '4', 'DGB200_PSPLIT_BM': '192'})     # pair split-K, cluster of 8
dense(600, 768, 1024, {'DGB200_CLUSTER': '4', 'DGB200_SPLITS': '1'})   # weight-multicast cluster
dense(256, 384, 512, majors='mm', fp32=True, c=True)                    # MN-major operands (wgrad form)
dense(200, 256, 512, {'DGB200_CLUSTER': '1', 'DGB200_SPLITS': '1'})    # single-CTA MMA, plain prologue

KNOBS = KNOBS + ('DGB200_SWAP', 'DGB200_GRID_TILES')
dense(300, 1000, 512, {'DGB200_SWAP': '1', 'DGB200_TMA_STORE': '1', 'DGB200_BLOCK_M': '96', 'DGB200_SPLITS': '1'})    # transposed output, staged
dense(120, 2000, 640, {'DGB200_SWAP': '1', 'DGB200_TMA_STORE': '0', 'DGB200_BLOCK_M': '48', 'DGB200_SPLITS': '1'})    # transposed output, direct
dense(700, 4096, 256, {'DGB200_GRID_TILES': '0', 'DGB200_SPLITS': '1'})                                              # persistent walk of a one-wave problem
for kk in KNOBS:
    os.environ.pop(kk, None)

# BF16 operands: dense nt (cluster split-K at this size), tn with FP32 accumulation (memory-side add), grouped, k-grouped, einsum
ab, bb = torch.randn((48, 1024), device='cuda', dtype=torch.bfloat16), torch.randn((384, 1024), device='cuda', dtype=torch.bfloat16)
db = torch.zeros((48, 384), device='cuda', dtype=torch.bfloat16)
dg.bf16_gemm_nt(ab, bb, db)
print('bf16 nt', _lib.last_config(), flush=True)
akm, bkn = torch.randn((512, 160), device='cuda', dtype=torch.bfloat16), torch.randn((512, 264), device='cuda', dtype=torch.bfloat16)
dacc = torch.zeros((160, 264), device='cuda', dtype=torch.float32)
dg.bf16_gemm_tn(akm, bkn, dacc, c=dacc)
gb = torch.randn((3, 256, 512), device='cuda', dtype=torch.bfloat16)
agr = torch.randn((3 * 128, 512), device='cuda', dtype=torch.bfloat16)
dgr = torch.zeros((3 * 128, 256), device='cuda', dtype=torch.bfloat16)
lay = torch.arange(3, device='cuda', dtype=torch.int32).repeat_interleave(128)
lay[100:128] = -1
dg.m_grouped_bf16_gemm_nt_contiguous(agr, gb, dgr, lay)
dg.m_grouped_bf16_gemm_nn_contiguous(agr, gb.transpose(1, 2).contiguous(), dgr, lay)
dmk = torch.zeros((3, 128, 256), device='cuda', dtype=torch.bfloat16)
dg.m_grouped_bf16_gemm_nt_masked(agr.view(3, 128, 512), gb, dmk, torch.tensor([7, 128, 0], device='cuda', dtype=torch.int32), 64)
ksb = [128, 0, 384]
akb, bkb = torch.randn((sum(ksb), 192), device='cuda', dtype=torch.bfloat16), torch.randn((sum(ksb), 136), device='cuda', dtype=torch.bfloat16)
dkb = torch.zeros((3, 192, 136), device='cuda', dtype=torch.float32)
dg.k_grouped_bf16_gemm_tn_contiguous(akb, bkb, dkb, ksb, torch.tensor(ksb, device='cuda', dtype=torch.int32), c=dkb)
xe, ye = torch.randn((40, 4, 256), device='cuda', dtype=torch.bfloat16), torch.randn((4, 128, 256), device='cuda', dtype=torch.bfloat16)
ze = torch.zeros((40, 4, 128), device='cuda', dtype=torch.bfloat16)
dg.einsum('bhr,hdr->bhd', xe, ye, ze)
ze2 = torch.zeros((40, 4, 256), device='cuda', dtype=torch.bfloat16)
dg.einsum('bhd,hdr->bhr', ze, ye, ze2)
abr, bbr = torch.randn((37, 136, 128), device='cuda', dtype=torch.bfloat16), torch.randn((37, 264, 128), device='cuda', dtype=torch.bfloat16)
dbr = torch.zeros((136, 264), device='cuda', dtype=torch.float32)
dg.einsum('bmk,bnk->mn', abr, bbr, dbr, c=dbr)
torch.cuda.synchronize()
print('bf16 family done', flush=True)

# skip_head_mid, bmm / einsum, quantiser
a = torch.randn((77, 384), device='cuda', dtype=torch.bfloat16)
b = torch.randn((768, 384), device='cuda', dtype=torch.bfloat16)
d = torch.zeros((77, 768 + 4 * 32), device='cuda', dtype=torch.bfloat16)
dg.fp8_gemm_nt_skip_head_mid(per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True), d, (64, 32, 128))
x = torch.randn((40, 4, 256), device='cuda', dtype=torch.bfloat16)
y = torch.randn((4, 384, 256), device='cuda', dtype=torch.bfloat16)
xq = per_token_cast_to_fp8(x.view(-1, 256), True)
xq = (xq[0].view(40, 4, 256), xq[1].view(40, 4, 2))
yq = [per_block_cast_to_fp8(y[i], True) for i in range(4)]
yq = (torch.stack([q[0] for q in yq]), torch.stack([q[1] for q in yq]))
z = torch.zeros((40, 4, 384), device='cuda', dtype=torch.bfloat16)
dg.fp8_einsum('bhr,hdr->bhd', xq, yq, z)
xc, yc = torch.randn((256, 2, 128), device='cuda', dtype=torch.bfloat16), torch.randn((256, 2, 256), device='cuda', dtype=torch.bfloat16)
xcq, ycq = per_channel_cast_to_fp8(xc.view(256, -1), True), per_channel_cast_to_fp8(yc.view(256, -1), True)
zz = torch.zeros((2, 128, 256), device='cuda', dtype=torch.float32)
dg.fp8_einsum('bhd,bhr->hdr', (xcq[0].view(256, 2, 128), xcq[1].view(2, 2, 128)), (ycq[0].view(256, 2, 256), ycq[1].view(2, 2, 256)), zz, zz,
              recipe=(1, 1, 128))
q, sf = dg.per_token_cast_to_fp8_packed(torch.randn((130, 640), device='cuda', dtype=torch.bfloat16))
torch.cuda.synchronize()
print('skip_head_mid / einsum / quantiser done', flush=True)

# grouped: EP dispatch (fused + overlap chain), psum GEMM, weighted top-k combine; contiguous with TMA store; masked; k-grouped
g, n, k = 4, 256, 512
w = torch.randn((g, n, k), device='cuda', dtype=torch.bfloat16)
qs = [per_block_cast_to_fp8(w[i], True) for i in range(g)]
wq = (torch.stack([q_[0] for q_ in qs]), torch.stack([q_[1] for q_ in qs]))
x = torch.randn((333, k), device='cuda', dtype=torch.bfloat16)
xq, sfp = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
ids = torch.randint(0, g, (333,), device='cuda')
ids[::9] = -1
buf = ep.EpBuffer(g, 333 * 2 + g * 128, k)
d = buf.output(n)
for ov in (False, True):
    _, r = ep.expert_sharded_grouped_gemm(xq, sfp, ids, wq, buf, d, overlap=ov)
    out = buf.combine(r.token_row, ids)
    torch.cuda.synchronize()
ids2 = torch.stack([torch.randperm(g, device='cuda')[:2] for _ in range(333)])
r = buf.dispatch(xq, sfp, ids2)
buf.grouped_gemm(wq, d, r.expected_m, overlap=False)
out = buf.combine(r.token_row, ids2, weights=torch.rand((333, 2), device='cuda'))
torch.cuda.synchronize()
buf.close()
print('ep done', flush=True)
os.environ['DGB200_TMA_STORE'] = '1'
layout = torch.arange(g, device='cuda', dtype=torch.int32).repeat_interleave(128)
a = torch.randn((g * 128, k), device='cuda', dtype=torch.bfloat16)
dc = torch.zeros((g * 128, n), device='cuda', dtype=torch.bfloat16)
dg.m_grouped_fp8_gemm_nt_contiguous(per_token_cast_to_fp8(a, True), wq, dc, layout)
os.environ.pop('DGB200_TMA_STORE')
a = torch.randn((g, 64, k), device='cuda', dtype=torch.bfloat16)
qa = [per_token_cast_to_fp8(a[i], True) for i in range(g)]
qa = (torch.stack([q_[0] for q_ in qa]), torch.stack([q_[1] for q_ in qa]))
dm = torch.zeros((g, 64, n), device='cuda', dtype=torch.bfloat16)
dg.m_grouped_fp8_gemm_nt_masked(qa, wq, dm, torch.tensor([5, 64, 0, 33], device='cuda', dtype=torch.int32), 32)
ks = [128, 0, 256]
ak, bk = torch.randn((sum(ks), 256), device='cuda', dtype=torch.bfloat16), torch.randn((sum(ks), 128), device='cuda', dtype=torch.bfloat16)
aq, bq = per_channel_cast_to_fp8(ak, True), per_channel_cast_to_fp8(bk, True)
dk = torch.zeros((3, 256, 128), device='cuda', dtype=torch.float32)
dg.k_grouped_fp8_gemm_tn_contiguous(aq, bq, dk, ks, torch.tensor(ks, device='cuda', dtype=torch.int32), c=dk)
torch.cuda.synchronize()
if 'LONG_K_LAST' in os.environ:
    dense(1, 2112, 7168)                                                # long K loop (14 k-blocks per slice): slow under racecheck
print('sanitize workload done', flush=True)
