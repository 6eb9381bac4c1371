"""Shared assertions of the GPU parity tests (the tolerance statement lives in tests/test_gemm_gpu.py's docstring)."""
import torch


def assert_close_to_oracle(d, want, what='', mag=None):
    """`mag`: magnitude that sets the BF16 rounding step (defaults to |want|; with C accumulation the product is rounded
    to BF16 BEFORE the add, so the step is that of |product| + |C|, not of the possibly cancelled sum)."""
    from deepgemm_b200.testing import calc_diff
    d, want = d.cpu(), want.cpu()
    mag = want.float().abs() if mag is None else mag.cpu().float()
    assert d.shape == want.shape, what
    if d.numel() == 0:
        return
    assert not torch.isnan(d.float()).any(), what
    assert calc_diff(d, want) < 1e-6, what
    if d.dtype == torch.bfloat16:
        mism = d != want
        assert mism.float().mean() <= 0.02, f'{what}: {int(mism.sum())} mismatches'
        err = (d.float() - want.float()).abs()
        tol = mag * 2.0 ** -7 + 1e-5 * mag.max()
        assert bool((err <= tol).all()), f'{what}: max excess {float((err - tol).max())}'
    else:
        scale = want.abs().max().clamp(min=1.0)
        assert ((d - want).abs().max() / scale) < 1e-5, what
