"""GPU tests of the peer-memory expert-parallel dispatch (csrc/ep_dispatch.cuh through the C ABI, deepgemm_b200/ep.py).

Single GPU: world size 1 exercises every phase of the fused dispatch kernel (rank / exchange / scatter / complete) and the
round-1 kernel chain of the dispatch || GEMM mode with the rank being its own peer; the result must be bit-identical to the torch re-layout `dispatch_local`. With >= 2 GPUs the torchrun script
tools/ep_check.py additionally checks the NVLink path against the NCCL all-to-all baseline and under a CUDA graph.
Reference context: the grouped GEMM inside expert parallelism, tests/test_mega_moe.py:148-205."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def dg():
    if not torch.cuda.is_available():
        pytest.skip('needs a GPU')
    import deepgemm_b200
    return deepgemm_b200


@pytest.mark.parametrize('id_dtype', [torch.int64, torch.int32])
@pytest.mark.parametrize('num_experts,t,k', [(8, 777, 1024), (32, 4096, 7168), (4, 5, 512)])
def test_peer_dispatch_world1_equals_torch_relayout(dg, num_experts, t, k, id_dtype):
    from deepgemm_b200 import ep
    from deepgemm_b200.utils import per_token_cast_to_fp8
    dev = torch.device('cuda', 0)
    gen = torch.Generator(device=dev).manual_seed(t)
    align = dg.get_mk_alignment_for_contiguous_layout()
    x = torch.randn((t, k), device=dev, dtype=torch.bfloat16, generator=gen)
    xq, sf = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    buf = ep.EpBuffer(num_experts, t + num_experts * align, k)
    try:
        for it in range(3):                                   # buffer reuse across epochs
            ids = torch.randint(0, num_experts if it < 2 else 2, (t,), device=dev, generator=gen).to(id_dtype)
            if it == 1:
                ids[::7] = -1                                 # unrouted tokens are skipped
            r = buf.dispatch(xq, sf, ids)
            torch.cuda.synchronize()
            keep = ids >= 0
            ref = ep.dispatch_local(xq[keep], sf[keep], ids[keep].long(), num_experts, align)
            m_al = ref.a.shape[0]
            assert buf.num_rows() == m_al and not buf.overflowed()
            assert torch.equal(r.psum_layout, ref.psum_layout)
            valid = ref.grouped_layout >= 0
            assert torch.equal(r.a[:m_al].view(torch.uint8)[valid], ref.a.view(torch.uint8)[valid])
            assert torch.equal(r.sfa[:m_al][valid], ref.sfa[valid])
            assert bool((r.token_row[~keep] == -1).all())
            rows = r.token_row[keep].long()
            assert torch.equal(r.a.view(torch.uint8)[rows], xq.view(torch.uint8)[keep])
            assert torch.equal(ref.grouped_layout[rows].long(), ids[keep].long())
    finally:
        buf.close()


def test_peer_dispatch_overflow_is_flagged_not_written(dg):
    from deepgemm_b200 import ep
    from deepgemm_b200.utils import per_token_cast_to_fp8
    dev = torch.device('cuda', 0)
    x = torch.randn((512, 512), device=dev, dtype=torch.bfloat16)
    xq, sf = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    buf = ep.EpBuffer(4, 256, 512)
    try:
        r = buf.dispatch(xq, sf, torch.zeros(512, dtype=torch.int64, device=dev))
        torch.cuda.synchronize()
        assert buf.overflowed()
        assert int((r.token_row >= 0).sum()) == buf.capacity == 256 and int(r.token_row.max()) == 255
    finally:
        buf.close()


@pytest.mark.parametrize('overlap', [False, True])
def test_expert_sharded_grouped_gemm_world1_matches_oracle(dg, overlap):
    from deepgemm_b200 import ep
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    from oracle import blockwise
    dev = torch.device('cuda', 0)
    g, n, k, t = 4, 256, 512, 333        # capacity 333 + 4 * 128 = 845 is rounded up to whole alignment units by EpBuffer
    gen = torch.Generator(device=dev).manual_seed(5)
    align = dg.get_mk_alignment_for_contiguous_layout()
    w = torch.randn((g, n, k), device=dev, dtype=torch.bfloat16, generator=gen)
    qs = [per_block_cast_to_fp8(w[i], True) for i in range(g)]
    wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
    x = torch.randn((t, k), device=dev, dtype=torch.bfloat16, generator=gen)
    xq, sf_packed = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    _, sf_fp32 = per_token_cast_to_fp8(x, True, 128)
    ids = torch.randint(0, g, (t,), device=dev, generator=gen)
    buf = ep.EpBuffer(g, t + g * align, k)
    try:
        d, r = ep.expert_sharded_grouped_gemm(xq, sf_packed, ids, wq, buf, overlap=overlap)
        torch.cuda.synchronize()
        rows = r.token_row.long().cpu()
        got = d.cpu()[rows].float()
        for e in range(g):
            sel = (ids == e).cpu()
            if not bool(sel.any()):
                continue
            want = blockwise.fp8_gemm_nt((xq.cpu()[sel], sf_fp32.cpu()[sel]), (wq[0][e].cpu(), wq[1][e].cpu())).float()
            mag = want.abs()
            assert bool(((got[sel] - want).abs() <= mag * 2.0 ** -7 + 1e-5 * mag.max()).all())
    finally:
        buf.close()


def test_combine_world1_returns_every_token_its_row(dg):
    from deepgemm_b200 import ep
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    dev = torch.device('cuda', 0)
    g, n, k, t = 8, 384, 512, 1000
    gen = torch.Generator(device=dev).manual_seed(9)
    align = dg.get_mk_alignment_for_contiguous_layout()
    w = torch.randn((g, n, k), device=dev, dtype=torch.bfloat16, generator=gen)
    qs = [per_block_cast_to_fp8(w[i], True) for i in range(g)]
    wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
    x = torch.randn((t, k), device=dev, dtype=torch.bfloat16, generator=gen)
    xq, sf_packed = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    ids = torch.randint(0, g, (t,), device=dev, generator=gen)
    ids[::11] = -1
    buf = ep.EpBuffer(g, t + g * align, k)
    try:
        d = buf.output(n)
        for _ in range(2):                                     # twice: epochs, buffer reuse
            d.fill_(float('nan'))
            _, r = ep.expert_sharded_grouped_gemm(xq, sf_packed, ids, wq, buf, d)
            out = buf.combine(r.token_row, ids)
            torch.cuda.synchronize()
            routed = ids >= 0
            assert torch.equal(out[routed], d[r.token_row[routed].long()])
            assert bool((out[~routed] == 0).all()) and not bool(torch.isnan(out.float()).any())
            # and the rows are the right answers: same as running every token through its expert densely
            for e in range(g):
                sel = ids == e
                if bool(sel.any()):
                    ref = torch.empty((int(sel.sum()), n), device=dev, dtype=torch.bfloat16)
                    _, sf_fp32 = per_token_cast_to_fp8(x[sel], True, 128)
                    dg.set_split_k(False)
                    try:
                        dg.fp8_gemm_nt((xq[sel].contiguous(), sf_fp32), (wq[0][e], wq[1][e]), ref)
                    finally:
                        dg.set_split_k(True)
                    assert torch.equal(out[sel], ref)
    finally:
        buf.close()


def _weighted_combine_reference(d_rows, weights, routed):
    """The kernel's arithmetic, spelled out: FP32 product, FP32 running sum in slot order, one BF16 rounding."""
    t, topk, n = d_rows.shape
    acc = torch.zeros((t, n), dtype=torch.float32, device=d_rows.device)
    for j in range(topk):
        term = weights[:, j:j + 1] * d_rows[:, j].float()
        acc = torch.where(routed[:, j:j + 1], acc + term, acc)
    return acc.to(torch.bfloat16)


@pytest.mark.parametrize('topk', [2, 8])
def test_topk_dispatch_and_weighted_combine_world1(dg, topk):
    """Top-k routing: every (token, slot) entry lands in its expert's segment; combine = FP32 weighted sum of the k expert
    outputs, bit-checked against a torch loop with the same operation order (tests/test_mega_moe.py:196-202 is the
    reference's baseline for this step)."""
    from deepgemm_b200 import ep
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    dev = torch.device('cuda', 0)
    g, n, k, t = 16, 256, 512, 700
    gen = torch.Generator(device=dev).manual_seed(40 + topk)
    align = dg.get_mk_alignment_for_contiguous_layout()
    w = torch.randn((g, n, k), device=dev, dtype=torch.bfloat16, generator=gen)
    qs = [per_block_cast_to_fp8(w[i], True) for i in range(g)]
    wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
    x = torch.randn((t, k), device=dev, dtype=torch.bfloat16, generator=gen)
    xq, sf_packed = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    _, sf_fp32 = per_token_cast_to_fp8(x, True, 128)
    ids = torch.stack([torch.randperm(g, device=dev, generator=gen)[:topk] for _ in range(t)])      # distinct experts per token
    ids[::13, 1] = -1                                                                               # some slots not routed
    weights = torch.rand((t, topk), device=dev, generator=gen)
    buf = ep.EpBuffer(g, t * topk + g * align, k)
    try:
        d = buf.output(n)
        for _ in range(2):
            d.fill_(float('nan'))
            r = buf.dispatch(xq, sf_packed, ids)
            buf.grouped_gemm(wq, d, r.expected_m, overlap=False)
            out = buf.combine(r.token_row, ids, weights=weights)
            torch.cuda.synchronize()
            assert not buf.overflowed()
            rows = r.token_row.view(t, topk)
            routed = ids >= 0
            assert bool((rows[~routed] == -1).all()) and bool((rows[routed] >= 0).all())
            # the dispatch put token t's bytes in every one of its rows
            flat = rows[routed].long()
            src = torch.arange(t, device=dev).unsqueeze(1).expand(t, topk)[routed]
            assert torch.equal(buf.a.view(torch.uint8)[flat], xq.view(torch.uint8)[src])
            assert flat.unique().numel() == flat.numel(), 'two entries share a row'
            # segment membership: row of an entry lies inside its expert's segment of the psum layout
            psum = r.psum_layout.long()
            starts = torch.cat([torch.zeros(1, dtype=torch.long, device=dev), (psum[:-1] + align - 1) // align * align])
            e = ids[routed].long()
            assert bool(((flat >= starts[e]) & (flat < psum[e])).all())
            # every row holds its expert's answer (dense kernel, split-K off: same bits)
            dg.set_split_k(False)
            try:
                for ex in (0, 7, 15):
                    sel = (ids == ex).any(dim=1)
                    if bool(sel.any()):
                        ref = torch.empty((int(sel.sum()), n), device=dev, dtype=torch.bfloat16)
                        dg.fp8_gemm_nt((xq[sel].contiguous(), sf_fp32[sel].contiguous()), (wq[0][ex], wq[1][ex]), ref)
                        slot = (ids[sel] == ex).float().argmax(dim=1)
                        assert torch.equal(d[rows[sel].gather(1, slot.unsqueeze(1)).squeeze(1).long()], ref)
            finally:
                dg.set_split_k(True)
            # the weighted reduce, bit for bit
            d_rows = d[rows.clamp(min=0).long()]
            want = _weighted_combine_reference(d_rows, weights, routed)
            assert torch.equal(out, want)
    finally:
        buf.close()


def test_overflowing_dispatch_followed_by_the_grouped_gemm_stays_in_bounds(dg):
    """with the psum ends clamped to `capacity` by an overflowing dispatch, the grouped GEMM must neither walk
    phantom tiles nor write past D: dropped tokens, no crash."""
    from deepgemm_b200 import ep
    from deepgemm_b200.utils import per_block_cast_to_fp8, per_token_cast_to_fp8
    dev = torch.device('cuda', 0)
    g, n, k, t = 4, 256, 512, 1200
    gen = torch.Generator(device=dev).manual_seed(11)
    w = torch.randn((g, n, k), device=dev, dtype=torch.bfloat16, generator=gen)
    qs = [per_block_cast_to_fp8(w[i], True) for i in range(g)]
    wq = (torch.stack([q[0] for q in qs]), torch.stack([q[1] for q in qs]))
    x = torch.randn((t, k), device=dev, dtype=torch.bfloat16, generator=gen)
    xq, sf_packed = per_token_cast_to_fp8(x, True, 128, use_packed_ue8m0=True)
    ids = torch.randint(0, g, (t,), device=dev, generator=gen)
    buf = ep.EpBuffer(g, 1000, k)                          # rounded up to 1024 rows: too small for 1200 tokens + padding
    try:
        guard = torch.full((buf.capacity + 256, n), 99.0, device=dev, dtype=torch.bfloat16)
        d = guard[:buf.capacity]
        r = buf.dispatch(xq, sf_packed, ids)
        buf.grouped_gemm(wq, d, r.expected_m, overlap=False)
        torch.cuda.synchronize()
        assert buf.overflowed()
        assert bool((guard[buf.capacity:] == 99.0).all())
        kept = r.token_row >= 0
        assert 0 < int(kept.sum()) < t and int(r.token_row.max()) < buf.capacity
    finally:
        buf.close()


def test_multi_gpu_peer_dispatch_under_torchrun(dg):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip('needs >= 2 GPUs on one node')
    n = 2 if n < 4 else 4
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', f'--nproc-per-node={n}', '--master-addr', '127.0.0.1',
           '--master-port', '29541', os.path.join(REPO, 'tools', 'ep_check.py')]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=REPO)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert res.stdout.count('ep check ok') == n
