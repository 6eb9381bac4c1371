"""k-grouped weight-gradient GEMMs (FP8 and BF16 operands, FP32 accumulate into D) beside the reference's kernels. Development tool."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference  # noqa: E402
import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200.testing import bench_kineto  # noqa: E402
from deepgemm_b200.utils import per_channel_cast_to_fp8  # noqa: E402

ref = import_reference()
for (g, m, n, ks) in [(4, 4096, 7168, [1024, 2048, 512, 4096]), (8, 4096, 4096, [4096] * 8), (4, 7168, 4096, [8192] * 4)]:
    sum_k = sum(ks)
    a = torch.randn((sum_k, m), device='cuda', dtype=torch.bfloat16)
    b = torch.randn((sum_k, n), device='cuda', dtype=torch.bfloat16)
    c = torch.randn((g, m, n), device='cuda', dtype=torch.float32)
    layout = torch.tensor(ks, device='cuda', dtype=torch.int32)
    flops = 2.0 * m * n * sum_k
    # BF16 operands
    d0, d1 = c.clone(), c.clone()
    ref.k_grouped_bf16_gemm_tn_contiguous(a, b, d0, ks, layout, c=d0)
    dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d1, ks, layout, c=d1)
    torch.cuda.synchronize()
    eq = bool(torch.equal(d0, d1))
    t_ref = bench_kineto(lambda: ref.k_grouped_bf16_gemm_tn_contiguous(a, b, d0, ks, layout, c=d0), 'sm100_bf16', num_tests=10)
    t_our = bench_kineto(lambda: dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d1, ks, layout, c=d1), 'fp8_gemm_kernel', num_tests=10)
    sweep = {}
    for bm in (os.environ.get('KG_SWEEP', '64,96,128,160,192,224').split(',')):
        os.environ['DGB200_BLOCK_M'] = bm
        try:
            sweep[bm] = round(bench_kineto(lambda: dg.k_grouped_bf16_gemm_tn_contiguous(a, b, d1, ks, layout, c=d1), 'fp8_gemm_kernel', num_tests=6) * 1e6, 1)
        except Exception as e:  # noqa: BLE001
            sweep[bm] = str(e)[:60]
        os.environ.pop('DGB200_BLOCK_M')
    print(json.dumps({'form': 'bf16 k_grouped_tn block_m sweep (us)', 'sweep': sweep}), flush=True)
    print(json.dumps({'form': 'bf16 k_grouped_tn', 'g': g, 'm': m, 'n': n, 'ks': ks, 'bitwise_equal': eq, 'ours_us': round(t_our * 1e6, 1),
                      'ref_us': round(t_ref * 1e6, 1), 'ours_tflops': round(flops / t_our / 1e12), 'ref_tflops': round(flops / t_ref / 1e12)}), flush=True)
    # FP8 operands, per-channel scales (tests/generators.py k-grouped case)
    a8, sfa = per_channel_cast_to_fp8(a, True)
    b8, sfb = per_channel_cast_to_fp8(b, True)
    d0, d1 = c.clone(), c.clone()
    try:
        ref.k_grouped_fp8_gemm_tn_contiguous((a8, sfa), (b8, sfb), d0, ks, layout, c=d0)
        dg.k_grouped_fp8_gemm_tn_contiguous((a8, sfa), (b8, sfb), d1, ks, layout, c=d1)
        torch.cuda.synchronize()
        eq = bool(torch.equal(d0, d1))
        t_ref = bench_kineto(lambda: ref.k_grouped_fp8_gemm_tn_contiguous((a8, sfa), (b8, sfb), d0, ks, layout, c=d0), 'sm100_fp8', num_tests=10)
        t_our = bench_kineto(lambda: dg.k_grouped_fp8_gemm_tn_contiguous((a8, sfa), (b8, sfb), d1, ks, layout, c=d1), 'fp8_gemm_kernel', num_tests=10)
        sweep = {}
        for bm in ('64', '128', '192'):
            os.environ['DGB200_BLOCK_M'] = bm
            try:
                sweep[bm] = round(bench_kineto(lambda: dg.k_grouped_fp8_gemm_tn_contiguous((a8, sfa), (b8, sfb), d1, ks, layout, c=d1), 'fp8_gemm_kernel', num_tests=6) * 1e6, 1)
            except Exception as e:  # noqa: BLE001
                sweep[bm] = str(e)[:60]
            os.environ.pop('DGB200_BLOCK_M')
        print(json.dumps({'form': 'fp8 k_grouped_tn block_m sweep (us)', 'sweep': sweep}), flush=True)
        print(json.dumps({'form': 'fp8 k_grouped_tn', 'g': g, 'm': m, 'n': n, 'ks': ks, 'bitwise_equal': eq, 'ours_us': round(t_our * 1e6, 1),
                          'ref_us': round(t_ref * 1e6, 1), 'ours_tflops': round(flops / t_our / 1e12), 'ref_tflops': round(flops / t_ref / 1e12)}), flush=True)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({'form': 'fp8 k_grouped_tn', 'error': str(e)[:300]}), flush=True)
    # dense FP32 accumulate (same epilogue)
    d0 = c[0].clone()
    a2, b2 = a[:ks[0]].t().contiguous(), b[:ks[0]].t().contiguous()
    t_acc = bench_kineto(lambda: dg.bf16_gemm_nt(a2, b2, d0, c=d0), 'fp8_gemm_kernel', num_tests=10)
    t_ref = bench_kineto(lambda: ref.bf16_gemm_nt(a2, b2, d0, c=d0), 'sm100_bf16', num_tests=10)
    print(json.dumps({'form': 'bf16 dense nt + fp32 C', 'm': m, 'n': n, 'k': ks[0], 'ours_us': round(t_acc * 1e6, 1), 'ref_us': round(t_ref * 1e6, 1)}), flush=True)
