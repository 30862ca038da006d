"""GPU: every C-ABI kernel against the torch-CPU restatement of its contract (tests/emu_backend.py,
itself pinned to the oracle by tests/test_host_logic_cpu.py).  fp32 tolerance 1e-4 relative to the
largest reference magnitude unless stated; index / layout ops are bit-exact."""
import pytest
import torch

from common import rel_err
from emu_backend import EmuBackend

pytestmark = pytest.mark.gpu

EMU = EmuBackend()
TOL = 1e-4


def _k(impl=0):
    from gansynth_b200.kernels import CudaBackend
    k = CudaBackend()
    k.impl = impl
    return k


def _rand(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


CONV_CASES = [
    # n, h, w, ci, co, ksize, stride
    (2, 8, 16, 8, 12, 3, 1),
    (2, 8, 16, 8, 12, 3, 2),
    (1, 6, 20, 4, 4, 3, 1),       # ragged tiles
    (1, 6, 20, 4, 4, 3, 2),
    (2, 32, 64, 32, 32, 3, 1),
    (2, 32, 64, 32, 64, 3, 2),
    (2, 16, 48, 64, 32, 3, 2),
    (1, 4, 32, 256, 256, 3, 1),
    (1, 4, 32, 128, 72, 3, 2),
    (2, 2, 16, 257, 256, 3, 1),   # stddev conv: naive path
    (2, 16, 32, 32, 2, 1, 1),     # to-RGB
    (2, 16, 32, 2, 32, 1, 1),     # from-RGB
    (3, 8, 24, 256, 2, 1, 1),     # to-RGB of a low-resolution colour block (growth path)
    (3, 8, 24, 2, 256, 1, 1),
    (1, 4, 8, 2, 64, 1, 1),
    (1, 4, 8, 128, 2, 1, 1),
    (1, 40, 24, 36, 20, 3, 1),    # K remainder (36 % 8 = 4)
    (2, 16, 32, 2, 16, 7, 2),     # pitch classifier: 7x7 stride-2 stem (SAME pads 2 / 3)
    (1, 12, 20, 3, 8, 5, 1),      # 5x5
    (2, 8, 16, 16, 32, 1, 2),     # 1x1 stride-2 projection shortcut
]


@pytest.mark.parametrize("case", CONV_CASES)
@pytest.mark.parametrize("wswap", [0, 1])
@pytest.mark.parametrize("impl", [0, 1, 2])
def test_conv_trio(case, wswap, impl):
    n, h, w, ci, co, ks, st = case
    if impl == 2 and (ks != 3 or ci % 4 or co % 4):
        pytest.skip("tiled kernel needs 3x3 and channels % 4 == 0")
    k = _k(impl)
    x = _rand(n, h, w, ci, seed=1)
    dy = _rand(n, h // st, w // st, co, seed=2)
    wt = _rand(ks, ks, co, ci, seed=3) if wswap else _rand(ks, ks, ci, co, seed=3)
    bias_c, bias_t = _rand(co, seed=4), _rand(ci, seed=5)
    alpha = 0.37
    for act in (0, 1):
        got = k.conv_c(x.cuda(), wt.cuda(), bias_c.cuda(), ks, st, wswap, alpha, act)
        want = EMU.conv_c(x, wt, bias_c, ks, st, wswap, alpha, act)
        assert got.shape == want.shape and rel_err(got, want) < TOL, "conv_c act=%d" % act
        got = k.conv_t(dy.cuda(), wt.cuda(), bias_t.cuda(), ks, st, wswap, alpha, act)
        want = EMU.conv_t(dy, wt, bias_t, ks, st, wswap, alpha, act)
        assert got.shape == want.shape and rel_err(got, want) < TOL, "conv_t act=%d" % act
    got = k.conv_c(x.cuda(), wt.cuda(), None, ks, st, wswap, alpha, 0)
    assert rel_err(got, EMU.conv_c(x, wt, None, ks, st, wswap, alpha, 0)) < TOL
    got = k.conv_w(x.cuda(), dy.cuda(), ks, st, wswap, alpha)
    want = EMU.conv_w(x, dy, ks, st, wswap, alpha)
    assert got.shape == want.shape and rel_err(got, want) < TOL, "conv_w"


def test_conv_transpose_abi_entry_points():
    """gs_conv2d_transpose_* against the oracle's conv2d_transpose (ops.py:250-280)."""
    from gansynth_b200 import _lib
    from oracle import ops as oops
    n, h, w, cin, f = 2, 4, 8, 8, 12
    x, var, b = _rand(n, cin, h, w, seed=1), _rand(3, 3, cin, f, seed=2), _rand(f, seed=3)
    want = oops.conv2d_transpose(x, var, b, (2, 2), 2.0)
    alpha = oops.he_constant(var.shape, 2.0)
    xd = x.permute(0, 2, 3, 1).contiguous().cuda()
    y = torch.empty(n, 2 * h, 2 * w, f, device="cuda")
    vd, bd = var.cuda(), b.cuda()
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("gs_conv2d_transpose_fwd", xd.data_ptr(), vd.data_ptr(), bd.data_ptr(), y.data_ptr(), n, h, w, cin, f, 3, 2,
              alpha, 0, 0, st)
    assert rel_err(y.permute(0, 3, 1, 2), want) < TOL
    # gradients of sum(y * g) w.r.t. x and var via autograd on the oracle
    g = _rand(n, f, 2 * h, 2 * w, seed=4)
    xr, vr = x.clone().requires_grad_(True), var.clone().requires_grad_(True)
    (oops.conv2d_transpose(xr, vr, None, (2, 2), 2.0) * g).sum().backward()
    gd = g.permute(0, 2, 3, 1).contiguous().cuda()
    dx = torch.empty(n, h, w, cin, device="cuda")
    dv = torch.empty(3, 3, cin, f, device="cuda")
    _lib.call("gs_conv2d_transpose_dgrad", gd.data_ptr(), vd.data_ptr(), dx.data_ptr(), n, h, w, cin, f, 3, 2, alpha, 0, st)
    _lib.call("gs_conv2d_transpose_wgrad", xd.data_ptr(), gd.data_ptr(), dv.data_ptr(), n, h, w, cin, f, 3, 2, alpha, 0, st)
    assert rel_err(dx.permute(0, 3, 1, 2), xr.grad) < TOL
    assert rel_err(dv, vr.grad) < TOL


@pytest.mark.parametrize("m,kk,n", [(8, 512, 8192), (8, 8192, 256), (8, 256, 61), (4, 96, 200), (12, 70, 33)])
def test_dense_trio(m, kk, n):
    k = _k()
    x, w, dy = _rand(m, kk, seed=1), _rand(kk, n, seed=2), _rand(m, n, seed=3)
    assert rel_err(k.dense_fwd(x.cuda(), w.cuda(), 0.5), EMU.dense_fwd(x, w, 0.5)) < TOL
    assert rel_err(k.dense_dgrad(dy.cuda(), w.cuda(), 0.5), EMU.dense_dgrad(dy, w, 0.5)) < TOL
    assert rel_err(k.dense_wgrad(x.cuda(), dy.cuda(), 0.5), EMU.dense_wgrad(x, dy, 0.5)) < TOL


def test_embedding():
    k = _k()
    table, idx, dy = _rand(61, 256, seed=1), torch.tensor([3, 60, 3, 0, 17, 17, 17, 5]), _rand(8, 256, seed=2)
    assert torch.equal(k.embedding_fwd(table.cuda(), idx.cuda(), 1.0).cpu(), EMU.embedding_fwd(table, idx, 1.0))
    assert rel_err(k.embedding_bwd(dy.cuda(), idx.cuda(), 61, 0.7), EMU.embedding_bwd(dy, idx, 61, 0.7)) < 1e-6


@pytest.mark.parametrize("shape", [(2, 8, 16, 32), (3, 5, 7, 61), (1, 2, 16, 256), (2, 4, 4, 2), (4, 512)])
def test_elementwise_and_pixel_norm(shape):
    k = _k()
    a, dy, u = _rand(*shape, seed=1), _rand(*shape, seed=2), _rand(*shape, seed=3)
    ac, dc, uc = a.cuda(), dy.cuda(), u.cuda()
    assert torch.equal(k.lrelu(ac).cpu(), EMU.lrelu(a))
    assert torch.equal(k.mask_mul(dc, ac).cpu(), EMU.mask_mul(dy, a))
    y = torch.tanh(a)
    assert rel_err(k.tanh_fwd(ac), y) < 1e-6
    assert rel_err(k.tanh_bwd(y.cuda(), dc), EMU.tanh_bwd(y, dy)) < 1e-6
    assert rel_err(k.tanh_bwd2(y.cuda(), dc, uc), EMU.tanh_bwd2(y, dy, u)) < 1e-6
    assert rel_err(k.axpby(ac, dc, 0.3, 0.7), EMU.axpby(a, dy, 0.3, 0.7)) < 1e-6
    assert rel_err(k.axpby(ac, None, 0.3, 0.0), EMU.axpby(a, None, 0.3, 0.0)) < 1e-6
    bias = _rand(shape[-1], seed=4)
    for act in (0, 1):
        assert rel_err(k.bias_act(ac, bias.cuda(), act), EMU.bias_act(a, bias, act)) < 1e-6
    assert torch.equal(k.row_broadcast(bias.cuda(), shape[:-1]).cpu(), EMU.row_broadcast(bias, shape[:-1]))
    assert rel_err(k.col_sum(ac), EMU.col_sum(a)) < 1e-5
    yk, rk = k.pn_fwd(ac, 1e-12)
    ye, re_ = EMU.pn_fwd(a, 1e-12)
    assert rel_err(yk, ye) < 1e-6 and rel_err(rk, re_) < 1e-6
    assert rel_err(k.pn_bwd(ac, rk, dc), EMU.pn_bwd(a, re_, dy)) < 1e-5
    assert rel_err(k.pn_bwd2(ac, rk, dc, uc), EMU.pn_bwd2(a, re_, dy, u)) < 1e-5
    # fused forms (fall back to the un-fused pair for channel counts the fused kernels do not cover)
    om, cm = k.mask_mul_colsum(dc, ac)
    oe, ce = EMU.mask_mul_colsum(dy, a)
    assert torch.equal(om.cpu(), oe) and rel_err(cm, ce.double()) < 1e-5
    for want in (False, True):
        zk, ck = k.pn_bwd_mask(ac, rk, dc, want)
        ze, ce = EMU.pn_bwd_mask(a, re_, dy, want)
        assert rel_err(zk, ze) < 1e-5
        assert (ck is None) == (not want)
        if want:
            assert rel_err(ck, ce.double()) < 1e-5
    gk, hk = k.pn_bwd_mask_second(ac, rk, dc, uc)
    ge, he = EMU.pn_bwd_mask_second(a, re_, dy, u)
    assert rel_err(gk, ge) < 1e-5 and rel_err(hk, he) < 1e-5


def test_fused_mask_colsum_full_size():
    """The top-resolution activation (8 x 128 x 1024 x 32): fused mask-multiply + bias gradient and fused
    pixel-norm backward + mask against the un-fused kernels."""
    k = _k()
    a, dy = _rand(8, 128, 1024, 32, seed=1).cuda(), _rand(8, 128, 1024, 32, seed=2).cuda()
    om, cm = k.mask_mul_colsum(dy, a)
    ref = k.mask_mul(dy, a)
    assert torch.equal(om, ref) and rel_err(cm, ref.double().sum((0, 1, 2))) < 1e-5
    _, r = k.pn_fwd(a, 1e-12)
    zk, ck = k.pn_bwd_mask(a, r, dy, True)
    zr = k.mask_mul(k.pn_bwd(a, r, dy), a)
    assert rel_err(zk, zr) < 1e-6 and rel_err(ck, zr.double().sum((0, 1, 2))) < 1e-5


def test_col_sum_large():
    k = _k()
    for rows, c in [(8 * 128 * 1024, 32), (8192, 256), (8, 8192), (1000, 61), (100000, 2)]:
        v = _rand(rows, c, seed=rows % 97)
        assert rel_err(k.col_sum(v.cuda()), v.double().sum(0)) < 1e-5


def test_batch_stddev():
    k = _k()
    for b, e in [(8, 2 * 16 * 256), (4, 4 * 4 * 256), (16, 100)]:
        x, u = _rand(b, e, seed=1), _rand(b, e, seed=2)
        df = _rand(b // 4, seed=3)
        assert rel_err(k.stddev_fwd(x.cuda(), 4, 1e-12), EMU.stddev_fwd(x, 4, 1e-12)) < 1e-5
        assert rel_err(k.stddev_bwd(x.cuda(), df.cuda(), 4, 1e-12), EMU.stddev_bwd(x, df, 4, 1e-12)) < 1e-5
        gk, qk = k.stddev_bwd2(x.cuda(), df.cuda(), u.cuda(), 4, 1e-12)
        ge, qe = EMU.stddev_bwd2(x, df, u, 4, 1e-12)
        assert rel_err(gk, ge) < 1e-5 and rel_err(qk, qe) < 1e-5


def test_resample_layout_rows():
    k = _k()
    x = _rand(2, 4, 8, 3, seed=1)
    assert torch.equal(k.upscale(x.cuda(), 2, 4, 1.0).cpu(), EMU.upscale(x, 2, 4, 1.0))
    big = _rand(2, 8, 32, 3, seed=2)
    assert rel_err(k.pool(big.cuda(), 2, 4, 0.125), EMU.pool(big, 2, 4, 0.125)) < 1e-6
    t = _rand(3, 5, 7, seed=3)
    assert torch.equal(k.transpose_inner(t.cuda()).cpu(), EMU.transpose_inner(t))
    a, b, s = _rand(8, 2 * 128 * 1024, seed=4), _rand(8, 2 * 128 * 1024, seed=5), _rand(8, seed=6)
    assert rel_err(k.row_dot(a.cuda(), b.cuda()), (a.double() * b.double()).sum(1)) < 1e-5
    assert rel_err(k.row_scale(a.cuda(), s.cuda()), EMU.row_scale(a, s)) < 1e-6


def test_adam_tf_semantics():
    k = _k()
    p, g = _rand(10001, seed=1), _rand(10001, seed=2)
    m, v = torch.zeros(10001), torch.zeros(10001)
    pc, mc, vc = p.clone().cuda(), m.clone().cuda(), v.clone().cuda()
    for t in (1, 2, 3):
        EMU.adam_step(p, g, m, v, 8e-4, 0.0, 0.99, 1e-8, t, 0.5)
        k.adam_step(pc, g.cuda(), mc, vc, 8e-4, 0.0, 0.99, 1e-8, t, 0.5)
    assert rel_err(pc, p) < 1e-5 and rel_err(vc, v) < 1e-5 and rel_err(mc, m) < 1e-5


def test_error_reporting():
    from gansynth_b200 import _lib
    with pytest.raises(_lib.GansynthLibraryError):
        _lib.call("gs_conv2d_fwd", 0, 0, 0, 0, 1, 8, 8, 4, 4, 4, 1, 0, 1.0, 0, 0, 0)   # ksize 4
    with pytest.raises(_lib.GansynthLibraryError):
        _k().lrelu(torch.zeros(4))   # CPU tensor: no fallback


def test_context_is_shared_by_the_host_threads_of_a_device():
    """The library binds its context (caller-owned workspace, split-weight cache) per host thread; PyTorch runs backward
    functions on its autograd thread, so every thread of the process binds the ONE context of the device: a second
    thread computes the same tensor-core convolution bit for bit, allocates no second workspace, and a cache reset
    from either thread reaches the shared cache."""
    import threading
    from gansynth_b200 import _lib
    from gansynth_b200.kernels import CudaBackend
    k = CudaBackend()
    g = torch.Generator().manual_seed(0)
    x, w = torch.randn(2, 16, 16, 32, generator=g).cuda(), torch.randn(3, 3, 32, 32, generator=g).cuda()
    main = k.conv_c(x, w, None, 3, 1, 0, 0.1, 0)
    torch.cuda.synchronize()
    before = len(_lib._contexts)
    out = {}

    def worker():
        torch.cuda.set_device(0)
        out["y"] = CudaBackend().conv_c(x, w, None, 3, 1, 0, 0.1, 0)
        CudaBackend().weight_cache_reset()
        torch.cuda.synchronize()

    t = threading.Thread(target=worker)
    t.start()
    t.join()
    assert len(_lib._contexts) == before == 1
    assert torch.equal(out["y"], main)


@pytest.mark.parametrize("shape,groups", [((2, 8, 16, 64), 32), ((3, 5, 7, 32), 4), ((1, 16, 64, 512), 32), ((2, 4, 4, 6), 3),
                                           ((2, 64, 128, 64), 32)])
@pytest.mark.parametrize("relu", [False, True])
def test_group_norm(shape, groups, relu):
    """gs_group_norm_fwd (ops.py:118-146, + the relu that follows every use in networks.py) against the emulation."""
    x = _rand(*shape, seed=1) * 2.0 + 0.5
    gamma, beta = 1.0 + 0.3 * _rand(shape[-1], seed=2), 0.2 * _rand(shape[-1], seed=3)
    got, stats = _k().group_norm(x.cuda(), gamma.cuda(), beta.cuda(), groups, 1e-12, relu)
    want, _ = EMU.group_norm(x.double(), gamma.double(), beta.double(), groups, 1e-12, relu)
    assert rel_err(got, want) < 1e-5
    # gradients (training the classifier): dx, dgamma, dbeta
    dy = _rand(*shape, seed=5)
    dx, dg, db = _k().group_norm_bwd(x.cuda(), got, dy.cuda(), stats, gamma.cuda(), groups, 1e-12, relu)
    ex, eg, eb = EMU.group_norm_bwd(x.double(), want, dy.double(), None, gamma.double(), groups, 1e-12, relu)
    assert rel_err(dx, ex) < 2e-4 and rel_err(dg, eg) < 2e-4 and rel_err(db, eb) < 2e-4, (rel_err(dx, ex), rel_err(dg, eg), rel_err(db, eb))


@pytest.mark.parametrize("shape,k,s", [((2, 16, 32, 64), 3, 2), ((1, 7, 9, 5), 3, 2), ((2, 8, 8, 16), 2, 2), ((1, 6, 10, 4), 3, 1)])
def test_max_pool_and_spatial_mean(shape, k, s):
    """gs_max_pool2d (ops.py:308-316, TF SAME: padding never wins) is an index op: bit-exact; gs_spatial_mean."""
    x = _rand(*shape, seed=4)
    y = _k().max_pool(x.cuda(), k, s)
    assert torch.equal(y.cpu(), EMU.max_pool(x, k, s))
    assert rel_err(_k().spatial_mean(x.cuda()), EMU.spatial_mean(x.double())) < 1e-6
    dy = _rand(*y.shape, seed=6)
    assert rel_err(_k().max_pool_bwd(x.cuda(), y, dy.cuda(), k, s), EMU.max_pool_bwd(x.double(), None, dy.double(), k, s)) < 1e-6
    # exact ties (a constant region, as the zero-padded head of a clip produces): the gradient goes to the FIRST maximum
    xt = x.clone()
    xt[:, : shape[1] // 2] = 0.25
    yt = _k().max_pool(xt.cuda(), k, s)
    assert rel_err(_k().max_pool_bwd(xt.cuda(), yt, dy.cuda(), k, s), EMU.max_pool_bwd(xt.double(), None, dy.double(), k, s)) < 1e-6
    dm = _rand(shape[0], shape[3], seed=7)
    assert rel_err(_k().spatial_mean_bwd(dm.cuda(), shape), EMU.spatial_mean_bwd(dm.double(), shape)) < 1e-6


@pytest.mark.parametrize("nesterov", [False, True])
def test_momentum_step(nesterov):
    """gs_momentum_step: tf.train.MomentumOptimizer with the weight decay folded in, two steps."""
    n = 10000
    p, g, wd = _rand(n, seed=1), _rand(n, seed=2), (torch.arange(n) % 3 == 0).float() * 1e-2
    pc, ac = p.clone().cuda(), torch.zeros(n).cuda()
    pe, ae = p.clone().double(), torch.zeros(n, dtype=torch.float64)
    for step in range(2):
        _k().momentum_step(pc, (g * (step + 1)).cuda(), ac, wd.cuda(), 0.05, 0.9, nesterov, 0.5)
        EMU.momentum_step(pe, (g * (step + 1)).double(), ae, wd.double(), 0.05, 0.9, nesterov, 0.5)
    assert rel_err(pc, pe) < 1e-6 and rel_err(ac, ae) < 1e-6
