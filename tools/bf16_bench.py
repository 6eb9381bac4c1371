"""BF16-operand GEMMs beside the reference's BF16 kernel (kernel time by the profiler, cold L2). Development tool."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference  # noqa: E402
import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200.testing import bench_kineto  # noqa: E402

ref = import_reference()
for (m, n, k) in [(64, 4096, 7168), (128, 2112, 7168), (512, 4096, 7168), (4096, 4096, 7168), (4096, 7168, 2048), (4096, 24576, 1536)]:
    a = torch.randn((m, k), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((n, k), device='cuda', dtype=torch.bfloat16)
    d0, d1 = torch.empty((m, n), device='cuda', dtype=torch.bfloat16), torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    ref.bf16_gemm_nt(a, b, d0)
    dg.bf16_gemm_nt(a, b, d1)
    torch.cuda.synchronize()
    t_ref = bench_kineto(lambda: ref.bf16_gemm_nt(a, b, d0), 'gemm', num_tests=10)
    t_our = bench_kineto(lambda: dg.bf16_gemm_nt(a, b, d1), 'fp8_gemm_kernel', num_tests=10)
    print(json.dumps({'m': m, 'n': n, 'k': k, 'bitwise_equal': bool(torch.equal(d0, d1)), 'ours_us': round(t_our * 1e6, 2), 'ref_us': round(t_ref * 1e6, 2),
                      'ours_tflops': round(2.0 * m * n * k / t_our / 1e12, 1), 'ref_tflops': round(2.0 * m * n * k / t_ref / 1e12, 1)}), flush=True)

# MN-major operands and the k-grouped weight-gradient form against the reference's kernels (same inputs)
for (m, n, k) in [(4096, 4096, 2048), (256, 7168, 4096)]:
    a = torch.randn((k, m), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((k, n), device='cuda', dtype=torch.bfloat16)
    d0, d1 = torch.empty((m, n), device='cuda', dtype=torch.bfloat16), torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
    ref.bf16_gemm_tn(a, b, d0)
    dg.bf16_gemm_tn(a, b, d1)
    torch.cuda.synchronize()
    t_ref = bench_kineto(lambda: ref.bf16_gemm_tn(a, b, d0), 'gemm', num_tests=10)
    t_our = bench_kineto(lambda: dg.bf16_gemm_tn(a, b, d1), 'fp8_gemm_kernel', num_tests=10)
    print(json.dumps({'form': 'tn', 'm': m, 'n': n, 'k': k, 'bitwise_equal': bool(torch.equal(d0, d1)), 'ours_us': round(t_our * 1e6, 2),
                      'ref_us': round(t_ref * 1e6, 2)}), flush=True)
g, m, n = 4, 4096, 7168
ks = [1024, 2048, 512, 4096]
a = torch.randn((sum(ks), m), device='cuda', dtype=torch.bfloat16)
b = torch.randn((sum(ks), n), device='cuda', dtype=torch.bfloat16)
c = torch.randn((g, m, n), device='cuda', dtype=torch.float32)
layout = torch.tensor(ks, device='cuda', dtype=torch.int32)
d0, d1 = c.clone(), c.clone()
ref.k_grouped_bf16_gemm_tn_contiguous(a, b, d0, ks, layout, c=d0)
dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d1, ks, layout, c=d1)
torch.cuda.synchronize()
t_ref = bench_kineto(lambda: ref.k_grouped_bf16_gemm_tn_contiguous(a, b, d0, ks, layout, c=d0), 'gemm', num_tests=10)
t_our = bench_kineto(lambda: dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d1, ks, layout, c=d1), 'fp8_gemm_kernel', num_tests=10)
print(json.dumps({'form': 'k_grouped_tn', 'g': g, 'm': m, 'n': n, 'ks': ks, 'bitwise_equal_after_1_call': None, 'ours_us': round(t_our * 1e6, 2),
                  'ref_us': round(t_ref * 1e6, 2)}), flush=True)
d0, d1 = c.clone(), c.clone()
ref.k_grouped_bf16_gemm_tn_contiguous(a, b, d0, ks, layout, c=d0)
dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d1, ks, layout, c=d1)
torch.cuda.synchronize()
print(json.dumps({'form': 'k_grouped_tn', 'bitwise_equal': bool(torch.equal(d0, d1)), 'max_abs_diff': float((d0 - d1).abs().max())}), flush=True)

# the batch-reduction einsum against the reference's dedicated kernel (tests/test_einsum.py:16-35)
for s_ in (129, 4096, 8192):
    for (m, n, k) in [(128, 384, 128), (256, 256, 256), (384, 128, 384)]:
        a = torch.randn((s_, m, k), device='cuda', dtype=torch.bfloat16)
        b = torch.randn((s_, n, k), device='cuda', dtype=torch.bfloat16)
        d0, d1 = torch.zeros((m, n), device='cuda'), torch.zeros((m, n), device='cuda')
        ref.einsum('bmk,bnk->mn', a, b, d0, c=d0)
        dg.einsum('bmk,bnk->mn', a, b, d1, c=d1)
        torch.cuda.synchronize()
        err = float((d0 - d1).abs().max() / d0.abs().max())
        t_ref = bench_kineto(lambda: ref.einsum('bmk,bnk->mn', a, b, d0, c=d0), 'bmn_bnk_mn_gemm_impl', num_tests=10)
        t_our = bench_kineto(lambda: dg.einsum('bmk,bnk->mn', a, b, d1, c=d1), 'fp8_gemm_kernel', num_tests=10)
        print(json.dumps({'form': 'bmk,bnk->mn', 's': s_, 'm': m, 'n': n, 'k': k, 'max_rel_diff_vs_ref': err, 'ours_us': round(t_our * 1e6, 1),
                          'ref_us': round(t_ref * 1e6, 1)}), flush=True)
