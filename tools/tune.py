"""Config sweeps on the big dense shapes (kernel-only, cold L2 via bench_kineto). usage: tune.py [big|small]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tools.bringup import import_reference, make_inputs  # noqa: E402
import deepgemm_b200 as dg  # noqa: E402
from deepgemm_b200 import _lib  # noqa: E402
from deepgemm_b200.testing import bench_kineto  # noqa: E402

KEYS = ('DGB200_M_BLOCKS', 'DGB200_BLOCK_M', 'DGB200_CLUSTER', 'DGB200_STAGES', 'DGB200_SWIZZLE_GROUP', 'DGB200_CSPLIT', 'DGB200_SPLITS', 'DGB200_PSPLIT',
        'DGB200_PSPLIT_BM', 'DGB200_TMA_STORE', 'DGB200_SWAP')


def setenv(**kw):
    for k in KEYS:
        os.environ.pop(k, None)
    for k, v in kw.items():
        if v is not None:
            os.environ['DGB200_' + k.upper()] = str(v)


def run(shapes, configs, with_ref=True):
    ref = import_reference() if with_ref else None
    for (m, n, k) in shapes:
        a, b, qa, qb = make_inputs(m, n, k)
        sfa = dg.transform_sf_into_required_layout(qa[1], m, k, (1, 128, 128), None, True)
        sfb = dg.transform_sf_into_required_layout(qb[1], n, k, (1, 128, 128), None, False)
        d = torch.empty((m, n), device='cuda', dtype=torch.bfloat16)
        rows = []
        if ref is not None:
            t = bench_kineto(lambda: ref.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d), 'gemm_', num_tests=10)
            rows.append(('ref', round(t * 1e6, 2)))
        for cfg in configs:
            setenv(**cfg)
            try:
                dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d)
                used = _lib.last_config()
                t = bench_kineto(lambda: dg.fp8_gemm_nt((qa[0], sfa), (qb[0], sfb), d), 'fp8_gemm_kernel', num_tests=10)
                rows.append((json.dumps(cfg), round(t * 1e6, 2), used['block_m'], used['num_stages'], used['cluster'], used['num_splits'],
                             'tma' if used['tma_store'] else 'direct', 'swap' if used['swap_ab'] else ''))
            except Exception as e:  # noqa: BLE001
                rows.append((json.dumps(cfg), 'error ' + str(e)[:80]))
        setenv()
        print(f'== {m}x{n}x{k}', flush=True)
        for r in rows:
            print('  ', *r, flush=True)


if __name__ == '__main__':
    mode = sys.argv[1] if len(sys.argv) > 1 else 'big'
    if mode == 'big':
        cfgs = [{}] + [dict(block_m=bm, swizzle_group=sg) for bm in (128, 160, 192, 208, 224, 240) for sg in (4, 8, 16)]
        cfgs += [dict(block_m=bm, stages=st) for bm in (224, 240) for st in (4, 5)]
        run([(4096, 7168, 2048), (4096, 4096, 7168)], cfgs)
    elif mode == 'store':
        # staged TMA-store epilogue on / off at the heuristics' own tile choice and at pinned heights
        cfgs = [{}, dict(tma_store=0), dict(tma_store=1)] + [dict(tma_store=ts, block_m=bm) for bm in (192, 208, 240) for ts in (0, 1)]
        run([(4096, 4096, 7168), (4096, 7168, 2048), (512, 4096, 7168), (512, 7168, 2048), (1024, 4096, 7168), (4096, 2112, 7168),
             (4096, 24576, 1536), (4096, 32768, 512), (4096, 7168, 16384)], cfgs)
    elif mode == 'mid':
        cfgs = [{}] + [dict(block_m=bm, csplit=0) for bm in (64, 96, 128, 192)]
        cfgs += [dict(psplit=2, psplit_bm=bm, csplit=0) for bm in (96, 128, 160, 192, 224)] + [dict(psplit=4, psplit_bm=bm, csplit=0) for bm in (128, 192)]
        run([(192, 4096, 7168), (256, 4096, 7168), (320, 4096, 7168), (384, 4096, 7168), (448, 4096, 7168), (512, 4096, 7168),
             (256, 7168, 2048), (512, 7168, 2048), (256, 2112, 7168), (384, 7168, 16384)], cfgs)
    elif mode == 'mid2':
        cfgs = [{}, dict(psplit=0)] + [dict(psplit=2, psplit_bm=bm, csplit=0) for bm in (64, 96, 128)]
        run([(160, 4096, 7168), (192, 4096, 7168), (224, 4096, 7168), (256, 4096, 7168), (320, 4096, 7168), (256, 2112, 7168), (384, 2112, 7168),
             (192, 7168, 16384), (256, 576, 7168), (256, 7168, 2048)], cfgs)
    elif mode == 'swap':
        # second orientation: tokens on the lanes, weight tiles of any multiple of 16 rows
        def auto_bn(n_, sms=148):
            return [bn for bn in sorted({-(-(-(-n_ // w)) // 16) * 16 for w in (sms, sms - 20, 2 * sms, 74, 3 * sms)}) if 16 <= bn <= 240]
        for shape in [(64, 7168, 2048), (128, 7168, 2048), (128, 24576, 1536), (64, 32768, 512), (128, 7168, 16384), (128, 2112, 7168),
                      (128, 576, 7168), (64, 4096, 7168), (128, 4096, 7168), (1, 7168, 2048), (1, 24576, 1536), (256, 4096, 7168),
                      (256, 7168, 2048), (512, 7168, 2048)]:
            cfgs = [dict(swap=0)] + [dict(swap=1, block_m=bn) for bn in auto_bn(shape[1])]
            if shape[0] > 128:
                cfgs += [dict(swap=1, block_m=bn) for bn in auto_bn(shape[1], 74)]
            run([shape], cfgs)
    elif mode == 'swap2':
        # second orientation with CTA pairs (256 tokens per tile): weight tiles cut so that all tiles make one / whole waves
        for shape, bns in [((512, 7168, 2048), (192, 208, 224, 240)), ((512, 4096, 7168), (112, 128)), ((4096, 7168, 2048), (208, 224, 240)),
                           ((4096, 7168, 16384), (224,)), ((1024, 7168, 2048), (208, 224)), ((384, 4096, 7168), (112, 128)),
                           ((448, 4096, 7168), (112, 128)), ((4096, 4096, 7168), (224, 240))]:
            run([shape], [dict(swap=0)] + [dict(swap=1, block_m=bn, tma_store=ts) for bn in bns for ts in (0, 1)])
    elif mode == 'ab_small':
        # default configuration only, on the latency-bound shapes: used to A/B two builds of the library (DGB200_LIB)
        run([(64, 4096, 7168), (128, 4096, 7168), (64, 7168, 2048), (128, 7168, 2048), (128, 24576, 1536), (64, 2112, 7168),
             (256, 4096, 7168), (512, 4096, 7168), (512, 7168, 2048)], [{}])
    elif mode == 'swap_exp':
        # the transposed-output kernel at the reference's own tiling (256 x 224) beside the default orientation at equal depth
        run([(4096, 7168, 2048)], [dict(swap=1, block_m=224, tma_store=1), dict(swap=0), dict(swap=0, block_m=224), dict(swap=0, block_m=224, stages=6),
                                   dict(swap=0, block_m=208, stages=6), dict(swap=1, block_m=224, tma_store=1, stages=5)])
    elif mode == 'swap3':
        # the transposed-output kernel with 64-column staged stores against the default orientation, shapes with N = 7168 / 24576 / 4096
        for shape, bns in [((4096, 7168, 2048), (224,)), ((4096, 7168, 16384), (224,)), ((1024, 7168, 2048), (224, 208)), ((2048, 7168, 2048), (224,)),
                           ((512, 7168, 2048), (224, 192)), ((4096, 4096, 7168), (224, 240, 192)), ((4096, 24576, 1536), (224, 240, 192)),
                           ((4096, 2112, 7168), (192, 224)), ((1024, 4096, 7168), (128, 224)), ((4096, 32768, 512), (224,)), ((2048, 4096, 7168), (224, 192))]:
            run([shape], [dict(swap=0)] + [dict(swap=1, block_m=bn, tma_store=1) for bn in bns])
    elif mode == 'store_exp':
        # default orientation, dominant shapes, staged vs direct epilogue (run against a build without the stores to see their cost)
        run([(4096, 4096, 7168), (4096, 7168, 2048)], [dict(swap=0), dict(swap=0, tma_store=0), dict(swap=0, tma_store=1, block_m=224)])
    elif mode == 'swap_small':
        # small M, many weight panels: single-CTA transposed-output tiles (tokens on the lanes) of a width that fills one wave
        for shape, bns in [((64, 7168, 2048), (48, 64, 96)), ((128, 7168, 2048), (48, 64, 96)), ((128, 24576, 1536), (160, 176, 192, 224)),
                           ((64, 24576, 1536), (160, 176, 192)), ((64, 32768, 512), (224, 240)), ((128, 7168, 16384), (48, 64))]:
            run([shape], [dict(swap=0)] + [dict(swap=1, block_m=bn, tma_store=ts) for bn in bns for ts in ((0, 1) if bn % 32 == 0 else (0,))])
    elif mode == 'swap_mid':
        # mid M at N = 4096, K = 7168: tokens on the lanes of a CTA pair (256 per tile), narrow weight tiles so that all SMs work
        for m_ in (224, 256, 320, 384, 448):
            run([(m_, 4096, 7168)], [dict(swap=0)] + [dict(swap=1, block_m=bn, tma_store=ts) for bn in (48, 56, 64, 96, 112) for ts in ((0, 1) if bn % 32 == 0 else (0,))])
    elif mode == 'mblocks':
        # dense, tensor-bound: the number of m-blocks (two heights each) behind the wave-balancing choice
        run([(4096, 4096, 7168)], [{}] + [dict(m_blocks=nb) for nb in (18, 19, 20, 21, 22, 23, 24, 25)] + [{}])
        run([(4096, 7168, 2048)], [{}] + [dict(m_blocks=nb) for nb in (18, 19, 20, 21, 22, 23)])
        run([(2048, 4096, 7168)], [{}] + [dict(m_blocks=nb) for nb in (9, 10, 11, 12, 13)])
    elif mode == 'mblocks2':
        run([(4096, 24576, 1536)], [{}] + [dict(m_blocks=nb) for nb in (18, 19, 20)])
        run([(3000, 4096, 7168)], [{}] + [dict(m_blocks=nb) for nb in (13, 14, 18)])
        run([(6144, 4096, 7168)], [{}] + [dict(m_blocks=nb) for nb in (27, 28, 32)])
        run([(8192, 4096, 7168)], [{}] + [dict(m_blocks=nb) for nb in (35, 36, 37)])
        run([(4096, 7168, 16384)], [{}] + [dict(m_blocks=nb) for nb in (18, 21)], with_ref=False)
    elif mode == 'tall':
        # the wave balancer's choice on tall dense problems (default configuration; + the direct epilogue where K is short)
        run([(4096, 4096, 7168), (4096, 7168, 2048), (3000, 4096, 7168), (2048, 4096, 7168), (6144, 4096, 7168), (8192, 4096, 7168),
             (4096, 2112, 7168), (1024, 4096, 7168)], [{}])
        run([(4096, 24576, 1536), (4096, 32768, 512), (4096, 7168, 16384)], [{}, dict(tma_store=0)])
    elif mode == 'small3':
        # short K with many weight panels at small M: single CTAs (192 / 56 independent tiles) vs pairs vs 2 single-CTA slices
        cfgs = [{}, dict(cluster=1, csplit=0), dict(cluster=1, csplit=0, block_m=64), dict(cluster=1, csplit=0, block_m=32), dict(csplit=2)]
        run([(64, 7168, 2048), (128, 7168, 2048), (128, 24576, 1536), (64, 32768, 512), (64, 24576, 1536), (128, 32768, 512)], cfgs)
    elif mode == 'small':
        cfgs = [{}, dict(csplit=0), dict(csplit=4), dict(csplit=2)] + [dict(csplit=0, block_m=bm) for bm in (16, 32, 64)]
        run([(1, 2112, 7168), (16, 4096, 7168), (32, 4096, 7168), (64, 4096, 7168), (96, 4096, 7168), (128, 4096, 7168), (192, 4096, 7168),
             (256, 4096, 7168), (64, 7168, 2048), (128, 7168, 2048), (64, 2112, 7168), (128, 24576, 1536), (64, 32768, 512), (128, 7168, 16384)], cfgs)
