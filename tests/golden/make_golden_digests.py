"""Golden DIGESTS of the reference's own SM100 kernel at full size (run on a B200 under gpurun; needs oracle/_ref).

For every case of tests/golden/cases.py (the reference's `enumerate_normal` shapes incl. the MN-major dgrad / wgrad
forms, BASELINE config 3 = contiguous G=256, config 4 = masked G=256) it runs the UNMODIFIED reference on the seeded
inputs and stores the SHA-256 of the output bytes (valid rows only for the grouped layouts). The outputs themselves are
up to 270 MB each, so only the digests are committed: gpurun_out/gpu_digests.json -> tests/golden/gpu_digests.json.

    python tests/golden/make_golden_digests.py            # all parts in parallel worker processes (JIT is CPU bound)
    python tests/golden/make_golden_digests.py --part 2/4 # one part
"""
import argparse
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, REPO)
sys.path.insert(0, HERE)


def run_part(part, parts, only=None):
    import torch
    os.environ.setdefault('DG_JIT_CACHE_DIR', '/tmp/dg_ref_cache')
    os.environ.setdefault('CUDA_HOME', '/usr/local/cuda')
    import cases
    from deepgemm_b200 import utils          # our quantisers (bit-identical to the reference's; inputs only)
    sys.path.insert(0, os.path.join(REPO, 'oracle', '_ref'))
    for k in [k for k in sys.modules if k == 'deep_gemm' or k.startswith('deep_gemm.')]:
        del sys.modules[k]
    import deep_gemm as ref
    assert 'oracle/_ref' in ref.__file__
    out = {}
    work = [('normal', c) for c in cases.normal_cases()] + [('contiguous', mm) for mm in (64, 128)] + \
           [('masked', mm) for mm in (16, 64, 96)] + [('bf16', shp) for shp in cases.BF16_CASES]
    for i, (kind, spec) in enumerate(work):
        if i % parts != part:
            continue
        if only and kind not in only:
            continue
        if kind == 'normal':
            qa, qb, c, d = cases.make_normal(spec, utils)
            if c is not None:
                d.copy_(c)
            ref.fp8_gemm_nt(qa, qb, d, c=d if c is not None else None)
            torch.cuda.synchronize()
            out[spec['name']] = cases.digest(d)
        elif kind == 'bf16':
            a, b = cases.make_bf16(*spec)
            d = torch.empty((spec[0], spec[1]), device='cuda', dtype=torch.bfloat16)
            ref.bf16_gemm_nt(a, b, d)
            torch.cuda.synchronize()
            out['bf16_%dx%dx%d' % spec] = cases.digest(d)
        elif kind == 'contiguous':
            p = cases.make_contiguous(spec, utils)
            d = torch.zeros((p['m'], p['b'][0].shape[1]), device='cuda', dtype=torch.bfloat16)
            ref.set_mk_alignment_for_contiguous_layout(p['alignment'])
            ref.m_grouped_fp8_gemm_nt_contiguous(p['a'], p['b'], d, p['layout'])
            torch.cuda.synchronize()
            out[f'contiguous_g256_m{spec}'] = cases.digest(d[p['valid']])
        else:
            p = cases.make_masked(spec, utils)
            g, m_max, _ = p['a'][0].shape
            d = torch.zeros((g, m_max, p['b'][0].shape[1]), device='cuda', dtype=torch.bfloat16)
            ref.m_grouped_fp8_gemm_nt_masked(p['a'], p['b'], d, p['masked_m'], p['expected_m'])
            torch.cuda.synchronize()
            out[f'masked_g256_m{spec}'] = cases.digest_masked(d, p['masked_m'])
        print(part, kind, spec if kind != 'normal' else spec['name'], flush=True)
        del d
        torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--part', default='')
    ap.add_argument('--parts', type=int, default=6)
    ap.add_argument('--only', default='', help='comma list of kinds (normal, contiguous, masked, bf16); merges into the committed file')
    args = ap.parse_args()
    os.makedirs(os.path.join(REPO, 'gpurun_out'), exist_ok=True)
    if args.part:
        part, parts = (int(x) for x in args.part.split('/'))
        res = run_part(part, parts, args.only.split(',') if args.only else None)
        with open(os.path.join(REPO, 'gpurun_out', f'gpu_digests.part{part}.json'), 'w') as f:
            json.dump(res, f)
        return
    extra = ['--only', args.only] if args.only else []
    procs = [subprocess.Popen([sys.executable, os.path.abspath(__file__), '--part', f'{i}/{args.parts}'] + extra) for i in range(args.parts)]
    rc = [p.wait() for p in procs]
    merged = {}
    committed = os.path.join(HERE, 'gpu_digests.json')
    if args.only and os.path.exists(committed):
        with open(committed) as f:
            merged.update({k: v for k, v in json.load(f).items() if k != '_meta'})
    for i in range(args.parts):
        path = os.path.join(REPO, 'gpurun_out', f'gpu_digests.part{i}.json')
        if os.path.exists(path):
            with open(path) as f:
                merged.update(json.load(f))
    merged['_meta'] = {'reference_kernel': 'sm100_fp8_fp4_gemm_1d1d_impl (oracle/_ref, unmodified)', 'worker_exit_codes': rc}
    with open(os.path.join(REPO, 'gpurun_out', 'gpu_digests.json'), 'w') as f:
        json.dump(merged, f, indent=1, sort_keys=True)
    print('wrote gpurun_out/gpu_digests.json with', len(merged) - 1, 'digests; exit codes', rc)


if __name__ == '__main__':
    main()
