"""GPU: tcgen05 / TMEM path.  The probe checks the descriptor encodings (SWIZZLE_NONE core-matrix
layouts, shifted window starts, K-major and MN-major operands) against a bf16-rounded torch matmul."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _bf16(x):
    return x.to(torch.bfloat16).to(torch.float64)


def _probe(a, b, k, n, shift, gstride, mode):
    from gansynth_b200 import _lib
    ad, bd = a.cuda().contiguous(), b.cuda().contiguous()
    d = torch.full((128, n), float("nan"), device="cuda")
    _lib.call("gs_tc_probe", ad.data_ptr(), bd.data_ptr(), d.data_ptr(), k, n, a.shape[0], b.shape[0], shift, gstride, mode,
              torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d.cpu().double()


@pytest.mark.parametrize("k,n,shift,gstride", [(16, 32, 0, 8), (64, 32, 0, 8), (32, 64, 0, 8), (32, 256, 0, 8),
                                               (32, 32, 1, 10), (32, 128, 11, 10), (64, 48, 22, 10)])
def test_probe_k_major(k, n, shift, gstride):
    g = torch.Generator().manual_seed(k * 1000 + n + shift)
    rows_a = 15 * gstride + 8 + shift + 3
    a = torch.randn(rows_a, k, generator=g)
    b = torch.randn(n, k, generator=g)
    got = _probe(a, b, k, n, shift, gstride, 0)
    pix = torch.tensor([(m // 8) * gstride + m % 8 + shift for m in range(128)])
    want = _bf16(a)[pix] @ _bf16(b).t()
    assert float((got - want).abs().max()) < 1e-3 * float(want.abs().max())


@pytest.mark.parametrize("k,n,shift,gstride", [(16, 32, 0, 8), (64, 64, 0, 8), (32, 256, 0, 8), (32, 32, 1, 10),
                                               (64, 128, 11, 10)])
def test_probe_mn_major(k, n, shift, gstride):
    g = torch.Generator().manual_seed(k * 1000 + n + shift + 7)
    rows = (k // 8 - 1) * gstride + 8 + shift + 2
    a = torch.randn(rows, 128, generator=g)
    b = torch.randn(rows, n, generator=g)
    got = _probe(a, b, k, n, shift, gstride, 1)
    pix = torch.tensor([(j // 8) * gstride + j % 8 + shift for j in range(k)])
    want = _bf16(a)[pix].t() @ _bf16(b)[pix]
    assert float((got - want).abs().max()) < 1e-3 * float(want.abs().max())
