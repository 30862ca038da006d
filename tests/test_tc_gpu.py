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
    _lib.probe_call("gs_tc_probe", ad.data_ptr(), bd.data_ptr(), d.data_ptr(), k, n, a.shape[0], b.shape[0], shift, gstride, mode,
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


# ------------------------------------------------------------------------------------------------
# tensor-core convolution forms (impl = 3) against the torch-CPU contract; bf16x3 products carry ~16
# mantissa bits, so the bound is 1e-4 of the largest output (measured ~1e-5).
from common import rel_err  # noqa: E402
from emu_backend import EmuBackend  # noqa: E402

EMU = EmuBackend()


def _rand(*shape, seed=0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed))


def _k3(impl=3):
    from gansynth_b200.kernels import CudaBackend
    k = CudaBackend()
    k.impl = impl
    return k


# impl 3: bf16 two-term split, three MMAs per K slice; truncating fp32 accumulation in TMEM (~1e-5)
TC_TOL = {3: 1e-4}


TC_FWD_CASES = [
    # n, h, w, ci, co, stride   (h, w = large side)
    (1, 16, 8, 32, 32, 1),
    (2, 16, 16, 32, 32, 1),
    (1, 32, 24, 64, 64, 1),
    (3, 48, 40, 96, 160, 1),
    (1, 16, 8, 256, 256, 1),
    (1, 128, 64, 32, 32, 1),
    (2, 32, 16, 32, 64, 2),
    (1, 64, 32, 64, 128, 2),
    (1, 32, 48, 128, 256, 2),
    # low-resolution 256-channel blocks: images interleaved row by row, output channels split over CTAs
    (8, 2, 16, 256, 256, 1),
    (8, 4, 32, 256, 256, 1),
    (3, 8, 64, 256, 256, 1),
    (8, 4, 32, 256, 256, 2),
    (8, 16, 128, 256, 256, 2),
    (5, 8, 16, 64, 32, 2),
    # thin wide layers: kw-stacked kernel (8 x 14 tiles; 128 = 9 * 14 + 2 and 136 = 9 * 14 + 10 end in partial tiles)
    (1, 8, 128, 32, 32, 1),
    (2, 16, 136, 32, 32, 1),
    (1, 8, 128, 64, 64, 1),
    (2, 32, 144, 64, 64, 1),
    (1, 16, 256, 64, 32, 1),
    (1, 8, 128, 32, 64, 1),
]


@pytest.mark.parametrize("case", TC_FWD_CASES)
@pytest.mark.parametrize("wswap", [0, 1])
@pytest.mark.parametrize("impl", [3])
def test_tc_gather_forms(case, wswap, impl):
    n, h, w, ci, co, st = case
    k = _k3(impl)
    x = _rand(n, h, w, ci, seed=1)
    wt = _rand(3, 3, co, ci, seed=3) if wswap else _rand(3, 3, ci, co, seed=3)
    bias = _rand(co, seed=4)
    for act, b in ((0, None), (1, bias)):
        got = k.conv_c(x.cuda(), wt.cuda(), None if b is None else b.cuda(), 3, st, wswap, 0.37, act)
        want = EMU.conv_c(x.double(), wt.double(), None if b is None else b.double(), 3, st, wswap, 0.37, act)
        assert got.shape == want.shape
        assert rel_err(got, want) < TC_TOL[impl], (case, wswap, act, rel_err(got, want))


TC_DGRAD_CASES = [
    # n, h, w (large side), ci (output channels of dgrad), co (contraction), stride
    (1, 16, 8, 32, 32, 1),
    (2, 32, 16, 64, 32, 1),
    (1, 16, 16, 256, 128, 1),
    (1, 32, 16, 32, 32, 2),
    (2, 64, 32, 32, 64, 2),
    (1, 32, 48, 128, 256, 2),
    (1, 64, 16, 64, 128, 2),
    (8, 2, 16, 256, 256, 1),
    (3, 4, 32, 256, 256, 1),
    (8, 8, 64, 256, 256, 1),
    (8, 4, 32, 256, 256, 2),
    (8, 16, 128, 256, 256, 2),
    (3, 8, 64, 64, 256, 2),
    (1, 8, 128, 32, 32, 1),
    (2, 16, 136, 64, 64, 1),
    (1, 32, 144, 32, 64, 1),
]


@pytest.mark.parametrize("case", TC_DGRAD_CASES)
@pytest.mark.parametrize("wswap", [0, 1])
@pytest.mark.parametrize("impl", [3])
def test_tc_transposed_forms(case, wswap, impl):
    n, h, w, ci, co, st = case
    k = _k3(impl)
    dy = _rand(n, h // st, w // st, co, seed=2)
    wt = _rand(3, 3, co, ci, seed=3) if wswap else _rand(3, 3, ci, co, seed=3)
    bias = _rand(ci, seed=5)
    for act, b in ((0, None), (1, bias)):
        got = k.conv_t(dy.cuda(), wt.cuda(), None if b is None else b.cuda(), 3, st, wswap, 0.37, act)
        want = EMU.conv_t(dy.double(), wt.double(), None if b is None else b.double(), 3, st, wswap, 0.37, act)
        assert got.shape == want.shape
        assert rel_err(got, want) < TC_TOL[impl], (case, wswap, act, rel_err(got, want))


def test_tc_full_size_layer_matches_fp32_tiled_kernel():
    """The top generator layer (8 x 128 x 1024 x 32 -> 32): tensor-core vs fp32 FFMA kernel."""
    from gansynth_b200.kernels import CudaBackend
    x = _rand(8, 128, 1024, 32, seed=1).cuda()
    wt = _rand(3, 3, 32, 32, seed=2).cuda()
    b = _rand(32, seed=3).cuda()
    k2, k3 = CudaBackend(), CudaBackend()
    k2.impl, k3.impl = 2, 3
    a = k2.conv_c(x, wt, b, 3, 1, 0, 0.0589, 1)
    c = k3.conv_c(x, wt, b, 3, 1, 0, 0.0589, 1)
    assert rel_err(c, a) < 1e-4


TC_WGRAD_CASES = [
    # n, h, w (large side), ci, co, stride
    (1, 16, 8, 32, 32, 1),
    (2, 32, 16, 32, 32, 1),
    (1, 16, 16, 64, 32, 1),
    (1, 16, 16, 32, 64, 1),
    (2, 16, 24, 64, 64, 1),
    (1, 8, 16, 128, 128, 1),
    (1, 4, 16, 256, 256, 1),
    (2, 2, 16, 256, 256, 1),
    (1, 16, 16, 256, 128, 1),
    (2, 32, 32, 32, 64, 2),
    (1, 32, 16, 64, 32, 2),
    (1, 16, 32, 128, 256, 2),
    (1, 8, 32, 256, 256, 2),
    (2, 128, 64, 32, 32, 1),
]


@pytest.mark.parametrize("case", TC_WGRAD_CASES)
@pytest.mark.parametrize("wswap", [0, 1])
def test_tc_filter_gradient(case, wswap):
    n, h, w, ci, co, st = case
    k = _k3(3)
    x = _rand(n, h, w, ci, seed=1)
    dy = _rand(n, h // st, w // st, co, seed=2)
    got = k.conv_w(x.cuda(), dy.cuda(), 3, st, wswap, 0.37)
    want = EMU.conv_w(x.double(), dy.double(), 3, st, wswap, 0.37)
    assert got.shape == want.shape
    assert rel_err(got, want) < 1e-4, (case, wswap, rel_err(got, want))


@pytest.mark.parametrize("case", TC_WGRAD_CASES + [(2, 64, 40, 96, 160, 1), (1, 8, 16, 8, 12, 1), (2, 8, 16, 8, 12, 2)])
@pytest.mark.parametrize("bias_of", ["dy", "x"])
@pytest.mark.parametrize("impl", [0, 4])
def test_filter_gradient_with_bias_gradient(case, bias_of, impl):
    """gs_conv2d_wgrad_ex: the bias gradient (column sum of either operand) taken inside the filter-gradient kernel --
    every pixel counted once although the kh jobs / N-stacked boxes of the tensor-core kernel overlap -- and the
    col_sum fallback of the fp32 kernels; dw itself unchanged."""
    n, h, w, ci, co, st = case
    k = _k3(impl)
    x = _rand(n, h, w, ci, seed=1)
    dy = _rand(n, h // st, w // st, co, seed=2)
    dw, db = k.conv_w(x.cuda(), dy.cuda(), 3, st, 0, 0.37, bias_of=bias_of)
    want_w, want_b = EMU.conv_w(x.double(), dy.double(), 3, st, 0, 0.37, bias_of=bias_of)
    assert dw.shape == want_w.shape and db.shape == want_b.shape
    assert rel_err(dw, want_w) < 1e-4, (case, rel_err(dw, want_w))
    assert rel_err(db, want_b) < 1e-5, (case, bias_of, rel_err(db, want_b))


# ------------------------------------------------------------------------------------------------
# fused epilogues (GS_EPI_MASK, GS_EPI_PIXEL_NORM): tensor-core tiles (both kernels, all three forms), shapes that
# fall back to convolution + elementwise kernel (split-K / split channel tiles / fp32 kernels), the (y, r) gradients
@pytest.mark.parametrize("case", [c for c in TC_FWD_CASES if c[:5] != (3, 48, 40, 96, 160)] + [(2, 8, 16, 8, 12, 1), (2, 8, 16, 8, 12, 2)])
@pytest.mark.parametrize("impl", [0, 4])
def test_mask_epilogue_gather(case, impl):
    n, h, w, ci, co, st = case
    k = _k3(impl)
    x, wt = _rand(n, h, w, ci, seed=1), _rand(3, 3, ci, co, seed=3)
    src = _rand(n, h // st, w // st, co, seed=6)
    got = k.conv_c(x.cuda(), wt.cuda(), None, 3, st, 0, 0.37, 0, mask_src=src.cuda())
    want = EMU.conv_c(x.double(), wt.double(), None, 3, st, 0, 0.37, 0, mask_src=src.double())
    assert got.shape == want.shape and rel_err(got, want) < 1e-4, (case, rel_err(got, want))


@pytest.mark.parametrize("case", TC_DGRAD_CASES + [(2, 8, 16, 8, 12, 1), (2, 8, 16, 8, 12, 2)])
@pytest.mark.parametrize("impl", [0, 4])
def test_mask_epilogue_transposed(case, impl):
    n, h, w, ci, co, st = case
    k = _k3(impl)
    dy, wt = _rand(n, h // st, w // st, co, seed=2), _rand(3, 3, ci, co, seed=3)
    src = _rand(n, h, w, ci, seed=6)
    got = k.conv_t(dy.cuda(), wt.cuda(), None, 3, st, 0, 0.37, 0, mask_src=src.cuda())
    want = EMU.conv_t(dy.double(), wt.double(), None, 3, st, 0, 0.37, 0, mask_src=src.double())
    assert got.shape == want.shape and rel_err(got, want) < 1e-4, (case, rel_err(got, want))


PN_CASES = [
    # form, n, h, w (large side), ci, co, stride
    ("c", 1, 8, 128, 32, 32, 1), ("c", 2, 16, 136, 32, 32, 1), ("c", 1, 8, 128, 64, 64, 1), ("c", 2, 32, 144, 64, 64, 1),
    ("c", 1, 16, 256, 64, 32, 1),                                   # kw-stacked kernel, one and two channel chunks
    ("c", 2, 16, 16, 32, 32, 1), ("c", 1, 32, 24, 64, 64, 1), ("c", 2, 32, 64, 128, 128, 1), ("c", 1, 16, 8, 256, 256, 1),
    ("c", 8, 2, 16, 256, 256, 1), ("c", 8, 4, 32, 256, 256, 1),    # split-K / split channel tiles: un-fused route
    ("t", 1, 32, 16, 32, 32, 2), ("t", 2, 64, 32, 32, 64, 2), ("t", 1, 32, 48, 128, 256, 2), ("t", 1, 64, 16, 64, 128, 2),
    ("t", 8, 4, 32, 256, 256, 2), ("t", 2, 128, 256, 32, 64, 2),
    ("c", 2, 8, 16, 8, 12, 1),                                      # fp32 kernels + elementwise pixel norm
]


@pytest.mark.parametrize("case", PN_CASES)
def test_pixel_norm_epilogue(case):
    form, n, h, w, ci, co, st = case
    k = _k3(0)
    bias = _rand(co if form == "c" else ci, seed=4)
    if form == "c":
        x, wt = _rand(n, h, w, ci, seed=1), _rand(3, 3, ci, co, seed=3)
    else:
        x, wt = _rand(n, h // st, w // st, co, seed=2), _rand(3, 3, ci, co, seed=3)
    y, r = k.conv_pn(x.cuda(), wt.cuda(), bias.cuda(), form, 3, st, 0, 0.37, 1e-8)
    ye, re_ = EMU.conv_pn(x.double(), wt.double(), bias.double(), form, 3, st, 0, 0.37, 1e-8)
    assert y.shape == ye.shape and r.shape == re_.shape
    assert rel_err(y, ye) < 1e-4 and rel_err(r, re_) < 1e-4, (case, rel_err(y, ye), rel_err(r, re_))


@pytest.mark.parametrize("shape", [(2, 8, 16, 32), (1, 4, 8, 64), (1, 2, 16, 256), (3, 4, 4, 128)])
def test_pixel_norm_y_form_gradients(shape):
    k = _k3(0)
    a, dy, u = _rand(*shape, seed=1), _rand(*shape, seed=2), _rand(*shape, seed=3)
    y, r = EMU.pn_fwd(a.double(), 1e-8)
    yc, rc = y.float().cuda(), r.float().cuda()
    for want_cs in (False, True):
        dz, cs = k.pn_bwd_mask_y(yc, rc, dy.cuda(), want_cs)
        ze, ce = EMU.pn_bwd_mask(a.double(), r, dy.double(), want_cs)
        assert rel_err(dz, ze) < 1e-5
        if want_cs:
            assert rel_err(cs, ce) < 1e-5
    ga, gdy = k.pn_bwd_mask_second_y(yc, rc, dy.cuda(), u.cuda())
    ge, he = EMU.pn_bwd_mask_second(a.double(), r, dy.double(), u.double())
    assert rel_err(ga, ge) < 1e-5 and rel_err(gdy, he) < 1e-5


def test_mask_epilogue_full_size_layer():
    """The discriminator's top layer gradient (8 x 128 x 1024, 32 -> 32) with the mask applied in the epilogue
    against the un-fused pair, bit for bit (same accumulators, same multiplication order)."""
    k = _k3(0)
    dy, wt = _rand(8, 128, 1024, 32, seed=1).cuda(), _rand(3, 3, 32, 32, seed=2).cuda()
    src = _rand(8, 128, 1024, 32, seed=3).cuda()
    fused = k.conv_t(dy, wt, None, 3, 1, 0, 0.0589, 0, mask_src=src)
    plain = k.conv_t(dy, wt, None, 3, 1, 0, 0.0589, 0)
    assert rel_err(fused, k.mask_mul(plain, src)) < 1e-6


def test_weight_cache_refresh_resplits_in_place():
    """gs_conv_weight_cache_refresh: after the parameters change, ONE batched launch re-splits every cached bf16 copy in
    its slot (all layouts: conv_tc 32- and 16-channel chunks, the kw-stacked layout, both weight orientations, flipped
    taps).  The refreshed cache must give exactly what a fresh split of the new values gives."""
    k = _k3(0)
    cases = [("c", 2, 16, 16, 64, 64, 1), ("t", 2, 16, 16, 64, 64, 1), ("c", 1, 16, 128, 32, 32, 1), ("t", 1, 16, 128, 32, 32, 1),
             ("c", 1, 32, 32, 32, 64, 2), ("t", 1, 32, 32, 32, 64, 2), ("c", 1, 8, 16, 256, 256, 1)]
    flat = (torch.randn(sum(9 * c[4] * c[5] for c in cases) + 64 * len(cases), generator=torch.Generator().manual_seed(5)) * 0.1).cuda()
    views, off = [], 0
    for c in cases:
        nel = 9 * c[4] * c[5]
        views.append(flat[off:off + nel].view(3, 3, c[4], c[5]))
        off += (nel + 63) // 64 * 64
    g = torch.Generator().manual_seed(6)
    inputs = [torch.randn(c[1], c[2] // (c[6] if c[0] == "t" else 1), c[3] // (c[6] if c[0] == "t" else 1), c[5] if c[0] == "t" else c[4],
                          generator=g).cuda() for c in cases]

    def run_all():
        return [(k.conv_c if c[0] == "c" else k.conv_t)(x, w, None, 3, c[6], 0, 0.3, 0) for c, x, w in zip(cases, inputs, views)]

    try:
        k.register_parameters([flat])            # weights inside `flat` are cacheable from here on
        run_all()                                # fills the cache
        with torch.no_grad():
            flat.mul_(-0.5).add_(0.01)           # an "optimiser update"
        k.weight_cache_refresh(flat)
        cached = run_all()
        k.register_parameters([])                # nothing cacheable: every call splits afresh
        fresh = run_all()
        for a, b, c in zip(cached, fresh, cases):
            if c[4] == 256:      # split-K layer: partial tiles meet through TMA reduce-add, whose order is not fixed
                assert rel_err(a, b) < 1e-6, c
            else:
                assert torch.equal(a, b), c
    finally:
        k.register_parameters([])
