"""GPU bring-up / sweep harness (development tool, run under gpurun). Not part of the product or of the tests.

  python tools/bringup.py dense   [--quick]    correctness of the dense kernel over configs (ours vs dequant-matmul)
  python tools/bringup.py ref                  ours vs the reference's SM100 kernel (oracle/_ref): bitwise + timing
  python tools/bringup.py sweep                timing sweep of block_m / stages on the headline shapes
Every line printed is JSON so results can be collected from gpurun_out/.
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def log(**kw):
    print(json.dumps(kw), flush=True)


def make_inputs(m, n, k, seed=0, device='cuda'):
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    g = torch.Generator(device=device).manual_seed(seed)
    a = torch.randn((m, k), device=device, dtype=torch.bfloat16, generator=g)
    b = torch.randn((n, k), device=device, dtype=torch.bfloat16, generator=g)
    return a, b, per_token_cast_to_fp8(a, True), per_block_cast_to_fp8(b, True)


def dequant_ref(qa, qb, m, n, k):
    """FP32 matmul of the exactly dequantised operands on the GPU (no TF32)."""
    torch.backends.cuda.matmul.allow_tf32 = False
    sfa = qa[1].repeat_interleave(128, 1)[:, :k]
    sfb = qb[1].repeat_interleave(128, 0)[:n].repeat_interleave(128, 1)[:, :k]
    return (qa[0].float() * sfa) @ (qb[0].float() * sfb).t()


def set_cfg(block_m=0, cluster=0, stages=0, splits=0):
    for name, v in (('DGB200_BLOCK_M', block_m), ('DGB200_CLUSTER', cluster), ('DGB200_STAGES', stages),
                    ('DGB200_SPLITS', splits)):
        if v:
            os.environ[name] = str(v)
        else:
            os.environ.pop(name, None)


def run_dense(quick):
    import deepgemm_b200 as dg
    from deepgemm_b200 import _lib
    from deepgemm_b200.testing import calc_diff
    shapes = [(128, 128, 128), (64, 256, 512), (128, 4096, 7168), (300, 2112, 1536), (4096, 4096, 7168), (1, 576, 7168)]
    cfgs = [(0, 1, 0), (0, 2, 0), (16, 2, 0), (32, 1, 3), (64, 2, 4), (128, 2, 0), (240, 2, 0), (208, 1, 2)]
    if quick:
        shapes, cfgs = shapes[:3], cfgs[:2]
    for (m, n, k) in shapes:
        a, b, qa, qb = make_inputs(m, n, k)
        ref = dequant_ref(qa, qb, m, n, k)
        for out_dtype in (torch.bfloat16, torch.float32):
            for (bm, cl, st) in cfgs:
                set_cfg(bm, cl, st)
                d = torch.full((m, n), float('nan'), device='cuda', dtype=out_dtype)
                t0 = time.time()
                try:
                    dg.fp8_gemm_nt(qa, qb, d)
                    torch.cuda.synchronize()
                except Exception as e:  # noqa: BLE001
                    log(test='dense', m=m, n=n, k=k, cfg=[bm, cl, st], error=str(e)[:300])
                    raise
                cfg = _lib.last_config()
                want = ref.to(out_dtype)
                exact = float((d == want).float().mean())
                log(test='dense', m=m, n=n, k=k, out=str(out_dtype), cfg=cfg, diff=calc_diff(d, ref),
                    max_abs=float((d.float() - ref).abs().max()), exact_frac=exact,
                    nan=int(torch.isnan(d.float()).sum()), ms=round((time.time() - t0) * 1e3, 2))
        # accumulate path
        set_cfg()
        for out_dtype in (torch.bfloat16, torch.float32):
            c = (torch.randn((m, n), device='cuda') * 32).to(out_dtype)
            d = c.clone()
            dg.fp8_gemm_nt(qa, qb, d, c=d)
            torch.cuda.synchronize()
            want = (ref.to(torch.bfloat16).float() + c.float()).to(out_dtype) if out_dtype == torch.bfloat16 else ref + c
            log(test='dense_acc', m=m, n=n, k=k, out=str(out_dtype), diff=calc_diff(d, want),
                exact_frac=float((d == want).float().mean()))
    set_cfg()


def import_reference():
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    ref_root = os.path.join(here, 'oracle', '_ref')
    os.environ.setdefault('DG_JIT_CACHE_DIR', '/tmp/dg_ref_cache')
    os.environ.setdefault('CUDA_HOME', '/usr/local/cuda')
    sys.path.insert(0, ref_root)
    for k in [k for k in sys.modules if k == 'deep_gemm' or k.startswith('deep_gemm.')]:
        del sys.modules[k]
    import deep_gemm as ref  # the UNMODIFIED reference
    assert ref_root in ref.__file__, ref.__file__
    return ref


def time_fn(fn, iters=20):
    from deepgemm_b200.testing import bench_events
    ts = bench_events(fn, num_warmups=3, num_tests=iters)
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def run_ref():
    ref = import_reference()
    import deepgemm_b200 as dg
    from deepgemm_b200.testing import bench_kineto
    set_cfg()
    os.environ['DG_PRINT_CONFIGS'] = '1'
    for (m, n, k) in [(128, 128, 128), (64, 4096, 7168), (128, 4096, 7168), (512, 4096, 7168), (4096, 4096, 7168),
                      (4096, 7168, 2048), (1, 2112, 7168)]:
        a, b, qa, qb = make_inputs(m, n, k)
        d_ref = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
        d_our = torch.empty_like(d_ref)
        t0 = time.time()
        ref.fp8_gemm_nt(qa, qb, d_ref)
        torch.cuda.synchronize()
        jit_s = time.time() - t0
        dg.fp8_gemm_nt(qa, qb, d_our)
        torch.cuda.synchronize()
        same = bool(torch.equal(d_ref, d_our))
        nmis = int((d_ref != d_our).sum())
        cfg_used = __import__('deepgemm_b200')._lib.last_config()
        # pre-packed SFs for both (kernel-only comparison)
        sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
        sfa_r = ref.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        sfb_r = ref.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
        pack_equal = bool(torch.equal(sfa, sfa_r) and torch.equal(sfb, sfb_r) and sfa.stride() == sfa_r.stride())
        t_ref = bench_kineto(lambda: ref.fp8_gemm_nt((qa[0], sfa_r), (qb[0], sfb_r), d_ref), 'gemm_')
        t_our = bench_kineto(lambda: dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d_our), 'fp8_gemm_kernel')
        e_ref, _ = time_fn(lambda: ref.fp8_gemm_nt(qa, qb, d_ref))
        e_our, _ = time_fn(lambda: dg.fp8_gemm_nt(qa, qb, d_our))
        fl = 2.0 * m * n * k
        log(test='ref_vs_ours', m=m, n=n, k=k, bitwise_equal=same, mismatches=nmis, pack_equal=pack_equal, cfg=cfg_used,
            ref_us=round(t_ref * 1e6, 2), our_us=round(t_our * 1e6, 2), ref_tflops=round(fl / t_ref / 1e12, 1) if t_ref else None,
            our_tflops=round(fl / t_our / 1e12, 1) if t_our else None, ref_e2e_us=round(e_ref * 1e6, 2),
            our_e2e_us=round(e_our * 1e6, 2), ref_jit_s=round(jit_s, 1))


def run_sweep():
    import deepgemm_b200 as dg
    from deepgemm_b200 import _lib
    for (m, n, k) in [(64, 4096, 7168), (128, 4096, 7168), (512, 4096, 7168), (4096, 4096, 7168)]:
        a, b, qa, qb = make_inputs(m, n, k)
        sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
        d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
        bms = [16, 32, 64, 128, 192, 224, 240]
        for bm in bms:
            if bm - 16 >= m and bm != 16:
                continue
            for cl in (2, 1):
                for st in (0, 4):
                    set_cfg(bm, cl, st)
                    try:
                        med, best = time_fn(lambda: dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d), iters=12)
                    except Exception as e:  # noqa: BLE001
                        log(test='sweep', m=m, n=n, k=k, cfg=[bm, cl, st], error=str(e)[:200])
                        continue
                    log(test='sweep', m=m, n=n, k=k, cfg=_lib.last_config(), med_us=round(med * 1e6, 2),
                        best_us=round(best * 1e6, 2), tflops=round(2.0 * m * n * k / med / 1e12, 1))
    set_cfg()


def make_grouped_weights(g, n, k, seed=0):
    """[G,N,K] FP8 weights + [G,N/128,K/128] scales, generated expert by expert to bound memory."""
    from deepgemm_b200.utils import per_block_cast_to_fp8
    gen = torch.Generator(device='cuda').manual_seed(seed)
    b = torch.empty((g, n, k), device='cuda', dtype=torch.float8_e4m3fn)
    sfb = torch.empty((g, (n + 127) // 128, (k + 127) // 128), device='cuda', dtype=torch.float32)
    for i in range(g):
        w = torch.randn((n, k), device='cuda', dtype=torch.bfloat16, generator=gen)
        b[i], sfb[i] = per_block_cast_to_fp8(w, True)
    return b, sfb


def run_grouped():
    """BASELINE configs 3 and 4 at full size, ours vs the reference kernel (kernel-only, cold L2)."""
    import random
    ref = import_reference()
    import deepgemm_b200 as dg
    from deepgemm_b200 import _lib
    from deepgemm_b200.testing import bench_kineto
    from deepgemm_b200.utils import per_token_cast_to_fp8
    set_cfg()
    random.seed(0)
    # ---- config 3: contiguous, 256 experts, N=4096, K=7168
    g, n, k = 256, 4096, 7168
    b, sfb = make_grouped_weights(g, n, k)
    sfb_p = dg.transform_sf_into_required_layout(sfb, n, k, (1, 128, 128), g, False)
    sfb_r = ref.transform_sf_into_required_layout(sfb, n, k, (1, 128, 128), g, False)
    log(test='grouped_sfb_pack_equal', equal=bool(torch.equal(sfb_p, sfb_r)))
    for mean_m in (64, 128, 256):
        for alignment in (128,):
            dg.set_mk_alignment_for_contiguous_layout(alignment)
            ref.set_mk_alignment_for_contiguous_layout(alignment)
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
            sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
            d_ref = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
            d_our = torch.empty_like(d_ref)
            ref.m_grouped_fp8_gemm_nt_contiguous((qa[0], sfa), (b, sfb_r), d_ref, layout)
            dg.m_grouped_fp8_gemm_nt_contiguous((qa[0], sfa), (b, sfb_p), d_our, layout)
            torch.cuda.synchronize()
            eq = bool(torch.equal(d_ref, d_our))
            cfg = _lib.last_config()
            t_ref = bench_kineto(lambda: ref.m_grouped_fp8_gemm_nt_contiguous((qa[0], sfa), (b, sfb_r), d_ref, layout), 'gemm_', num_tests=5)
            t_our = bench_kineto(lambda: dg.m_grouped_fp8_gemm_nt_contiguous((qa[0], sfa), (b, sfb_p), d_our, layout), 'fp8_gemm_kernel', num_tests=5)
            valid = sum(ms)
            byts = m * k + g * n * k + m * n * 2
            log(test='contiguous', groups=g, mean_m=mean_m, m=m, valid_m=valid, n=n, k=k, alignment=alignment, cfg=cfg,
                bitwise_equal=eq, ref_us=round(t_ref * 1e6, 1), our_us=round(t_our * 1e6, 1),
                ref_tflops=round(2.0 * valid * n * k / t_ref / 1e12, 1), our_tflops=round(2.0 * valid * n * k / t_our / 1e12, 1),
                ref_gbs=round(byts / t_ref / 1e9), our_gbs=round(byts / t_our / 1e9))
            del a, qa, sfa, d_ref, d_our, layout
    del b, sfb, sfb_p, sfb_r
    torch.cuda.empty_cache()
    # ---- config 4: masked, 256 experts, M_max=128, N=7168, K=2048
    g, m_max, n, k = 256, 128, 7168, 2048
    b, sfb = make_grouped_weights(g, n, k, seed=1)
    sfb_p = dg.transform_sf_into_required_layout(sfb, n, k, (1, 128, 128), g, False)
    a = torch.randn((g, m_max, k), device='cuda', dtype=torch.bfloat16)
    qs = [per_token_cast_to_fp8(a[i], True) for i in range(g)]
    qa = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
    sfa = dg.transform_sf_into_required_layout(qa[1], m_max, k, (1, 128, 128), g, True)
    for mean_m in (16, 64, 96):
        masked = torch.tensor([min(m_max, int(mean_m * random.uniform(0.7, 1.3))) for _ in range(g)], device='cuda', dtype=torch.int32)
        expected_m = int(1.2 * mean_m)
        d_ref = torch.zeros((g, m_max, n), device='cuda', dtype=torch.bfloat16)
        d_our = torch.zeros_like(d_ref)
        ref.m_grouped_fp8_gemm_nt_masked((qa[0], sfa), (b, sfb_p), d_ref, masked, expected_m)
        dg.m_grouped_fp8_gemm_nt_masked((qa[0], sfa), (b, sfb_p), d_our, masked, expected_m)
        torch.cuda.synchronize()
        cfg = _lib.last_config()
        eq = all(bool(torch.equal(d_ref[i, :mm], d_our[i, :mm])) for i, mm in enumerate(masked.tolist()))
        t_ref = bench_kineto(lambda: ref.m_grouped_fp8_gemm_nt_masked((qa[0], sfa), (b, sfb_p), d_ref, masked, expected_m), 'gemm_', num_tests=5)
        t_our = bench_kineto(lambda: dg.m_grouped_fp8_gemm_nt_masked((qa[0], sfa), (b, sfb_p), d_our, masked, expected_m), 'fp8_gemm_kernel', num_tests=5)
        # CUDA-graph replay of ours
        graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream()
        with torch.cuda.stream(side):
            with torch.cuda.graph(graph, stream=side):
                dg.m_grouped_fp8_gemm_nt_masked((qa[0], sfa), (b, sfb_p), d_our, masked, expected_m)
        t_graph, _ = time_fn(lambda: graph.replay(), iters=8)
        valid = int(masked.sum())
        byts = valid * k + g * n * k + valid * n * 2
        log(test='masked', groups=g, mean_m=mean_m, valid_m=valid, n=n, k=k, cfg=cfg, valid_rows_bitwise_equal=eq,
            ref_us=round(t_ref * 1e6, 1), our_us=round(t_our * 1e6, 1), our_graph_us=round(t_graph * 1e6, 1),
            ref_tflops=round(2.0 * valid * n * k / t_ref / 1e12, 1), our_tflops=round(2.0 * valid * n * k / t_our / 1e12, 1),
            ref_gbs=round(byts / t_ref / 1e9), our_gbs=round(byts / t_our / 1e9))


def run_mcast():
    """Weight-multicast clusters (4 / 8 CTAs) vs plain pairs: bitwise check + kernel time."""
    import deepgemm_b200 as dg
    from deepgemm_b200 import _lib
    from deepgemm_b200.testing import bench_kineto
    for (m, n, k, bms) in [(64, 4096, 7168, (16, 32, 64)), (128, 4096, 7168, (32, 64, 128)), (512, 4096, 7168, (64, 128)),
                           (4096, 4096, 7168, (128, 192, 240)), (4096, 7168, 2048, (128, 240)), (1024, 2112, 7168, (128, 240))]:
        a, b, qa, qb = make_inputs(m, n, k)
        sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
        base = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
        set_cfg(bms[-1], 2, 0, 1)
        dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), base)
        for bm in bms:
            for cl in (2, 4, 8):
                if cl // 2 * bm > max(m, bm) * 2 and cl > 2 and (m + bm - 1) // bm < cl // 2:
                    continue
                set_cfg(bm, cl, 0, 1)
                d = torch.full((m, n), float('nan'), device='cuda', dtype=torch.bfloat16)
                try:
                    dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d)
                    torch.cuda.synchronize()
                    t = bench_kineto(lambda: dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d), 'fp8_gemm_kernel', num_tests=10)
                except Exception as e:  # noqa: BLE001
                    log(test='mcast', m=m, n=n, k=k, bm=bm, cluster=cl, error=str(e)[:200])
                    raise
                log(test='mcast', m=m, n=n, k=k, cfg=_lib.last_config(), equal=bool(torch.equal(d, base)),
                    us=round(t * 1e6, 2), tflops=round(2.0 * m * n * k / t / 1e12, 1))
    set_cfg()


if __name__ == '__main__':
    mode = sys.argv[1]
    torch.manual_seed(0)
    log(mode=mode, device=torch.cuda.get_device_name(0), sms=torch.cuda.get_device_properties(0).multi_processor_count)
    if mode == 'dense':
        run_dense('--quick' in sys.argv)
    elif mode == 'ref':
        run_ref()
    elif mode == 'sweep':
        run_sweep()
    elif mode == 'grouped':
        run_grouped()
    elif mode == 'mcast':
        run_mcast()
    log(mode=mode, done=True)
