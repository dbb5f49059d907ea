"""GPU: each accx kernel, called through the C ABI (ctypes), against a plain torch fp32
restatement of the same operator on the same seeded inputs.  Tolerances: fp32 storage
rtol 1e-3 (atol 1e-5 * max), bf16 storage rtol 2e-2 (atol 1e-2 * max)."""
import pytest
import torch
import torch.nn.functional as F

from helpers import close

pytestmark = pytest.mark.gpu

DEV = "cuda"
# the torch restatement must be real fp32 (cuDNN/cuBLAS default to TF32 for convolutions)
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
TOL = {torch.float32: (1e-3, 1e-5), torch.bfloat16: (2e-2, 1e-2)}


def E():
    from accx import engine
    return engine


def lrelu(x):
    return torch.where(x > 0, x, 0.01 * x)


def mk_lazy(shape, dtype, act, seed):
    g = torch.Generator(device="cpu").manual_seed(seed)
    x = torch.randn(shape, generator=g).to(DEV).to(dtype)
    C = shape[-1]
    if act == 0:
        return E().Lazy(x), x.float()
    s = (torch.rand(C, generator=g) + 0.5).to(DEV)
    t = (torch.randn(C, generator=g) * 0.3).to(DEV)
    a = x.float() * s + t
    if act == 2:
        a = lrelu(a)
    return E().Lazy(x, s, t, act), a


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,N,act", [(3, 9, 0), (9, 3, 2), (45, 3, 2), (32, 96, 0), (96, 32, 2), (480, 64, 1),
                                     (128, 136, 2), (16, 256, 2)])
def test_pw_fwd_single(dtype, K, N, act):
    e = E()
    B, H, W = 2, 12, 20
    L, a = mk_lazy((B, H, W, K), dtype, act, 1)
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    stats = torch.zeros(2 * N, device=DEV)
    y = e.conv([e.Op(L, K, e.WV(w, 0, K, 1))], N, (B, H, W), bias=b, stats=stats)
    ref = a @ w.t() + b
    rt, at = TOL[dtype]
    close(y.float(), ref, rt, at, "pw_fwd")
    close(stats[:N], ref.sum((0, 1, 2)), 1e-3, 1e-3, "stats sum")
    close(stats[N:], (ref * ref).sum((0, 1, 2)), 1e-3 if dtype == torch.float32 else 2e-2, 1e-3, "stats sumsq")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pw_fwd_multi_operand_strided_weights_and_adds(dtype):
    """two operands reading interleaved weight columns (K index 2c+j) + two upsample-adds + fp32 output"""
    e = E()
    B, H, W, C, N = 2, 8, 16, 24, 40
    L0, a0 = mk_lazy((B, H, W, C), dtype, 2, 3)
    L1, a1 = mk_lazy((B, H, W, C), dtype, 0, 4)
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(N, 2 * C, generator=g) / C ** 0.5).to(DEV)
    add1 = torch.randn(B, H // 2, W // 2, N, generator=g).to(DEV)
    add2 = torch.randn(B, H // 4, W // 4, N, generator=g).to(DEV)
    y = e.conv([e.Op(L0, C, e.WV(w, 0, 2 * C, 2)), e.Op(L1, C, e.WV(w, 1, 2 * C, 2))], N, (B, H, W),
               adds=[(add1, 1), (add2, 2)], out_dtype=e.F32)
    ref = a0 @ w[:, 0::2].t() + a1 @ w[:, 1::2].t()
    ref = ref + add1.repeat_interleave(2, 1).repeat_interleave(2, 2) + add2.repeat_interleave(4, 1).repeat_interleave(4, 2)
    assert y.dtype == torch.float32
    rt, at = TOL[dtype]
    close(y, ref, rt, at, "pw_fwd multi")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pw_fwd_channel_offset_operands(dtype):
    """operands that are column slices of a wider matrix (the pooled avg|max buffer) and sliced output"""
    e = E()
    B, H, W, C, N = 1, 4, 8, 16, 8
    L, a = mk_lazy((B, H, W, 2 * C), dtype, 0, 6)
    g = torch.Generator().manual_seed(7)
    w = torch.randn(N, 5 * C, generator=g).to(DEV) / 4
    out = torch.zeros(B, H, W, 3 * N, device=DEV, dtype=torch.float32)
    e.conv([e.Op(L, C, e.WV(w, 1, 5 * C, 5), 0), e.Op(L, C, e.WV(w, 3, 5 * C, 5), C)], N, (B, H, W), out=out, out_coff=N)
    ref = a[..., :C] @ w[:, 1::5].t() + a[..., C:] @ w[:, 3::5].t()
    rt, at = TOL[dtype]
    close(out[..., N:2 * N], ref, rt, at, "sliced out")
    assert float(out[..., :N].abs().max()) == 0 and float(out[..., 2 * N:].abs().max()) == 0


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C", [8, 20])
def test_dense3x3_as_nine_shifted_operands(dtype, C):
    e = E()
    B, H, W = 2, 9, 12
    L, a = mk_lazy((B, H, W, C), dtype, 0, 8)
    g = torch.Generator().manual_seed(9)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    b = torch.randn(C, generator=g).to(DEV)
    ops = [e.Op(L, C, e.WV(w, ky * 3 + kx, C * 9, 9), 0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    y = e.conv(ops, C, (B, H, W), bias=b)
    ref = F.conv2d(a.permute(0, 3, 1, 2), w, b, padding=1).permute(0, 2, 3, 1)
    rt, at = TOL[dtype]
    close(y.float(), ref, rt, at, "conv3x3")
    # input gradient = the transposed-tap contraction
    dy = torch.randn(B, H, W, C, generator=g).to(DEV).to(dtype)
    opsT = [e.Op(e.Lazy(dy), C, e.WV(w, ky * 3 + kx, 9, C * 9), 0, 1 - ky, 1 - kx) for ky in range(3) for kx in range(3)]
    dx = e.conv(opsT, C, (B, H, W))
    a_ = a.clone().requires_grad_(True)
    F.conv2d(a_.permute(0, 3, 1, 2), w, b, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    close(dx.float(), a_.grad, rt, at, "conv3x3 dgrad")
    # weight gradient, tap by tap
    gw = torch.zeros_like(w)
    for op in ops:
        e.wgrad(op, dy, C, (B, H, W), gw)
    w_ = w.clone().requires_grad_(True)
    F.conv2d(a.permute(0, 3, 1, 2), w_, b, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    close(gw, w_.grad, rt, at, "conv3x3 wgrad")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("K,N,act", [(3, 9, 0), (96, 32, 2), (40, 136, 1)])
def test_pw_wgrad(dtype, K, N, act):
    e = E()
    B, H, W = 2, 10, 14
    L, a = mk_lazy((B, H, W, K), dtype, act, 10)
    g = torch.Generator().manual_seed(11)
    dy = torch.randn(B, H, W, N, generator=g).to(DEV).to(dtype)
    w = torch.zeros(N, K, device=DEV)
    gw = torch.zeros(N, K, device=DEV)
    e.wgrad(e.Op(L, K, e.WV(w, 0, K, 1)), dy, N, (B, H, W), gw)
    ref = torch.einsum("bhwn,bhwk->nk", dy.float(), a)
    rt, at = TOL[dtype]
    close(gw, ref, rt, at, "pw_wgrad")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,H,W", [(9, 7, 5), (96, 12, 10), (24, 8, 8), (272, 4, 6), (64, 32, 64), (384, 24, 24),
                                   (160, 16, 16), (96, 48, 96), (200, 9, 33)])
def test_dw3x3(dtype, C, H, W):
    e = E()
    B = 2
    L, a = mk_lazy((B, H, W, C), dtype, 2, 12)
    g = torch.Generator().manual_seed(13)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(DEV)
    b = torch.randn(C, generator=g).to(DEV)
    stats = torch.zeros(2 * C, device=DEV)
    y = e.dw_fwd(L, w, b, stats)
    a_ = a.clone().requires_grad_(True)
    w_ = w.clone().requires_grad_(True)
    ref = F.conv2d(a_.permute(0, 3, 1, 2), w_, b, padding=1, groups=C).permute(0, 2, 3, 1)
    rt, at = TOL[dtype]
    close(y.float(), ref, rt, at, "dw fwd")
    close(stats[:C], ref.sum((0, 1, 2)), 1e-3, 2e-3, "dw stats")
    dy = torch.randn(B, H, W, C, generator=g).to(DEV).to(dtype)
    ref.backward(dy.float())
    dx = e.dw_fwd(e.Lazy(dy), w, None, None, flip=True)
    close(dx.float(), a_.grad, rt, at, "dw dgrad")
    gw = torch.zeros_like(w)
    e.dw_wgrad(L, dy, gw)
    close(gw, w_.grad, rt, at, "dw wgrad")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("k", [2, 3, 5])
def test_hanc_pools_and_unpool(dtype, k):
    e = E()
    B, H, W, C = 2, 16, 32, 24
    L, a = mk_lazy((B, H, W, C), dtype, 2, 14)
    pools = e.hanc_pools(L, k)
    a_ = a.clone().requires_grad_(True)
    an = a_.permute(0, 3, 1, 2)
    rt, at = TOL[dtype]
    total = 0
    g = torch.Generator().manual_seed(15)
    da = torch.zeros(B, H, W, C, device=DEV, dtype=dtype)
    for l in range(1, k):
        s = 1 << l
        avg = F.avg_pool2d(an, s).permute(0, 2, 3, 1)
        mx = F.max_pool2d(an, s).permute(0, 2, 3, 1)
        close(pools[l - 1][..., :C].float(), avg, rt, at, f"avg{s}")
        close(pools[l - 1][..., C:].float(), mx, rt, at, f"max{s}")
        dp = torch.randn(B, H >> l, W >> l, 2 * C, generator=g).to(DEV)
        total = total + (avg * dp[..., :C]).sum() + (mx * dp[..., C:]).sum()
        e.hanc_unpool_bwd(L, l, dp, da, accumulate=(l > 1))
    total.backward()
    close(da.float(), a_.grad, rt, at, "unpool bwd")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,H,W", [(96, 12, 10), (24, 8, 8), (64, 32, 64), (384, 24, 24), (200, 9, 33)])
def test_dw3x3_dgrad_fused_with_bn_reduce(dtype, C, H, W):
    """accx_dw3x3_dgrad_bnred == accx_dw3x3_fwd(flip) followed by accx_bn_bwd_reduce"""
    e = E()
    B = 2
    L1, _ = mk_lazy((B, H, W, C), dtype, 2, 21)
    g = torch.Generator().manual_seed(22)
    L1.mean = (torch.randn(C, generator=g) * 0.2).to(DEV)
    L1.rstd = (torch.rand(C, generator=g) + 0.5).to(DEV)
    w = (torch.randn(C, 1, 3, 3, generator=g) / 3).to(DEV)
    dy = torch.randn(B, H, W, C, generator=g).to(DEV).to(dtype)
    ar = e.Arena(DEV)
    assert e.dw_dgrad_bnred_ok(L1, dy)
    da, sums = e.dw_dgrad_bnred(L1, dy, w, ar)
    ref = e.dw_fwd(e.Lazy(dy), w, None, None, flip=True)
    assert torch.equal(da, ref)
    sums_ref = torch.zeros(2 * C, device=DEV)
    e._call("accx_bn_bwd_reduce", e.dt(ref), B * H * W, C, e.ptr(L1.y), e.ptr(L1.scale), e.ptr(L1.shift), 2, e.ptr(L1.mean),
            e.ptr(L1.rstd), e.ptr(ref), e.ptr(sums_ref), e.stream())
    rt = 1e-3 if dtype == torch.float32 else 1e-2
    close(sums, sums_ref, rt, rt, "fused dw dgrad bn-backward sums")


@pytest.mark.parametrize("k", [2, 3])
@pytest.mark.parametrize("C,H,W", [(24, 16, 32), (96, 8, 12), (272, 4, 8)])
def test_hanc_unpool_fused_with_bn_reduce_matches_separate_kernels(k, C, H, W):
    """accx_hanc_unpool_bnred == accx_hanc_unpool_bwd per level followed by accx_bn_bwd_reduce (bf16);
    ties inside a window (frequent in bf16) must be routed identically: first maximum in row-major order."""
    e = E()
    B = 2
    g = torch.Generator().manual_seed(31)
    y = torch.randn(B, H, W, C, generator=g).to(DEV).to(torch.bfloat16)
    y[:, ::2, 1::2, :] = y[:, ::2, ::2, :]            # force exact ties inside every 2x2 / 4x4 window
    scale = (torch.rand(C, generator=g) + 0.5).to(DEV)
    shift = (torch.randn(C, generator=g) * 0.3).to(DEV)
    mean = (torch.randn(C, generator=g) * 0.2).to(DEV)
    rstd = (torch.rand(C, generator=g) + 0.5).to(DEV)
    L = e.Lazy(y, scale, shift, 2, mean, rstd, None)
    da0 = torch.randn(B, H, W, C, generator=g).to(DEV).to(torch.bfloat16)
    dps = [torch.randn(B, H >> l, W >> l, 2 * C, generator=g).to(DEV) for l in range(1, k)]
    ar = e.Arena(DEV)
    # separate kernels (fp32 copy of da so that the reference carries no intermediate bf16 rounding)
    ref = da0.float().clone()
    Lf = e.Lazy(y.float(), scale, shift, 2, mean, rstd, None)
    for l, dp in enumerate(dps, start=1):
        e.hanc_unpool_bwd(Lf, l, dp, ref, accumulate=True)
    sums_ref = torch.zeros(2 * C, device=DEV)
    e._call("accx_bn_bwd_reduce", e.F32, B * H * W, C, e.ptr(Lf.y), e.ptr(scale), e.ptr(shift), 2, e.ptr(mean), e.ptr(rstd),
            e.ptr(ref), e.ptr(sums_ref), e.stream())
    da = da0.clone()
    sums = e.hanc_unpool_bnred(L, dps, da, ar)
    close(da.float(), ref, 1e-2, 1e-2, "fused unpool da")
    close(sums, sums_ref, 2e-3, 2e-3, "fused bn-backward sums")


def test_unpool_routes_ties_to_first_maximum():
    """ATen MaxPool2d sends the gradient to the first maximal element in row-major window order."""
    e = E()
    x = torch.zeros(1, 4, 4, 8, device=DEV)
    x[0, 0, 2, :] = 1.0
    x[0, 1, 0, :] = 1.0        # tie inside the single 4x4 window: (0,2) comes first in row-major order
    dp = torch.zeros(1, 1, 1, 16, device=DEV)
    dp[..., 8:] = 1.0
    da = torch.zeros_like(x)
    e.hanc_unpool_bwd(e.Lazy(x), 2, dp, da, accumulate=False)
    assert float(da[0, 0, 2, 0]) == 1.0 and float(da[0, 1, 0, 0]) == 0.0
    xr = x.permute(0, 3, 1, 2).clone().requires_grad_(True)
    F.max_pool2d(xr, 4).sum().backward()
    assert torch.equal(xr.grad.permute(0, 2, 3, 1), da)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_pool_sum_and_upsample_add(dtype):
    e = E()
    B, H, W, C = 2, 8, 16, 40
    g = torch.Generator().manual_seed(16)
    x = torch.randn(B, H, W, C, generator=g).to(DEV).to(dtype)
    rt, at = TOL[dtype]
    for l in (0, 1, 3):
        ref = F.avg_pool2d(x.float().permute(0, 3, 1, 2), 1 << l).permute(0, 2, 3, 1)
        close(e.pool_sum(x, l, 1.0 / 4 ** l).float(), ref, rt, at, f"avg pool {l}")
        close(e.pool_sum(x, l, 1.0, torch.float32), ref * 4 ** l, rt, at, f"sum pool {l}")
    src = torch.randn(B, H // 4, W // 4, C, generator=g).to(DEV).to(dtype)
    dst = torch.randn(B, H, W, C, generator=g).to(DEV).to(dtype)
    ref = dst.float() + 0.5 * src.float().repeat_interleave(4, 1).repeat_interleave(4, 2)
    e.upsample_add(src, dst, 2, 0.5, accumulate=True)
    close(dst.float(), ref, rt, at, "upsample_add")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("C,act", [(9, 2), (96, 2), (64, 1), (4352, 2), (8, 2), (768, 2), (1536, 1), (520, 2), (3, 1)])
def test_batchnorm_forward_stats_and_backward(dtype, C, act):
    e = E()
    B, H, W = 2, 6, 10
    g = torch.Generator().manual_seed(17)
    y = (torch.randn(B, H, W, C, generator=g) * 2 + 0.5).to(DEV).to(dtype)
    bn = torch.nn.BatchNorm2d(C).to(DEV)
    with torch.no_grad():
        bn.weight.copy_(torch.rand(C, generator=g) + 0.5)
        bn.bias.copy_(torch.randn(C, generator=g) * 0.2)
    ref_bn = torch.nn.BatchNorm2d(C).to(DEV)
    ref_bn.load_state_dict(bn.state_dict())
    ar = e.Arena(DEV)
    stats = ar.take(2 * C)
    e.materialize(e.Lazy(y), stats=stats, stats_only=True)
    L = e.bn_lazy(y, stats, bn, act, ar, True)
    out = e.materialize(L)
    y_ = y.float().clone().requires_grad_(True)
    ref = ref_bn(y_.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    if act == 2:
        ref = lrelu(ref)
    rt, at = TOL[dtype]
    close(out.float(), ref, rt, at, "bn fwd")
    close(bn.running_mean, ref_bn.running_mean, 1e-3, 1e-4, "running_mean")
    close(bn.running_var, ref_bn.running_var, 1e-3, 1e-4, "running_var")
    assert int(bn.num_batches_tracked) == 1
    da = torch.randn(B, H, W, C, generator=g).to(DEV).to(dtype)
    ref.backward(da.float())
    grads = {}
    dy = e.bn_bwd(L, da.clone(), grads, ar)
    close(dy.float(), y_.grad, rt, 5 * at, "bn bwd dx")
    close(grads[id(bn.weight)], ref_bn.weight.grad, rt, 10 * at, "dgamma")
    close(grads[id(bn.bias)], ref_bn.bias.grad, rt, 10 * at, "dbeta")


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layout_roundtrip(dtype):
    e = E()
    x = torch.randn(3, 5, 7, 9, device=DEV)
    n = e.input_to_nhwc(x, dtype)
    close(n.float(), x.permute(0, 2, 3, 1), 1e-2 if dtype == torch.bfloat16 else 0, 1e-2 if dtype == torch.bfloat16 else 0, "to nhwc")
    v = e.to_nchw_view(n)
    assert v.shape == x.shape
    assert e.to_nhwc(v).data_ptr() == n.data_ptr()        # channels_last in -> no copy
    assert e.to_nhwc(x.to(dtype)).shape == n.shape
