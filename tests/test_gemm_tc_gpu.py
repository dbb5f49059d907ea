"""GPU: the tcgen05/TMEM contraction (accx_pw_fwd_tc) against a torch fp32 restatement computed
from the same bf16 inputs.  bf16 operands, fp32 accumulation: rtol 2e-2, atol 1e-2 * max."""
import pytest
import torch
import torch.nn.functional as F

from helpers import close
from test_kernels_gpu import DEV, E, mk_lazy, lrelu

pytestmark = pytest.mark.gpu
RT, AT = 2e-2, 1e-2


def bf(x):
    return x.to(torch.bfloat16).float()


@pytest.fixture(autouse=True)
def tc_on():
    e = E()
    old = e.TC
    e.TC = True
    yield
    e.TC = old


@pytest.mark.parametrize("P_shape", [(2, 12, 20), (1, 16, 8), (3, 7, 11)])
@pytest.mark.parametrize("K,N,act", [(32, 96, 0), (96, 32, 2), (480, 64, 1), (128, 136, 2), (16, 256, 2), (64, 272, 2),
                                     (8, 8, 0), (1088, 128, 2), (24, 40, 2)])
def test_tc_single_operand(P_shape, K, N, act):
    e = E()
    B, H, W = P_shape
    L, a = mk_lazy((B, H, W, K), torch.bfloat16, act, 1)
    g = torch.Generator().manual_seed(2)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    b = torch.randn(N, generator=g).to(DEV)
    stats = torch.zeros(2 * N, device=DEV)
    n0 = e.LAUNCHES_EXTRA[0]
    y = e.conv([e.Op(L, K, e.WV(w, 0, K, 1))], N, (B, H, W), bias=b, stats=stats)
    assert e.LAUNCHES_EXTRA[0] == n0 + 1, "tensor-core path was not taken"
    ref = bf(a) @ bf(w).t() + b
    close(y.float(), ref, RT, AT, "tc pw_fwd")
    close(stats[:N], ref.sum((0, 1, 2)), 2e-2, 2e-2, "tc stats sum")
    close(stats[N:], (ref * ref).sum((0, 1, 2)), 2e-2, 2e-2, "tc stats sumsq")


def test_tc_matches_simt_path_bitwise_inputs():
    """same call through both paths (tensor cores vs CUDA cores)"""
    e = E()
    B, H, W, K, N = 2, 24, 24, 192, 64
    L, a = mk_lazy((B, H, W, K), torch.bfloat16, 2, 3)
    w = (torch.randn(N, K, generator=torch.Generator().manual_seed(4)) / K ** 0.5).to(DEV)
    y_tc = e.conv([e.Op(L, K, e.WV(w, 0, K, 1))], N, (B, H, W), out_dtype=e.F32)
    e.TC = False
    y_simt = e.conv([e.Op(L, K, e.WV(w, 0, K, 1))], N, (B, H, W), out_dtype=e.F32)
    close(y_tc, y_simt, 1e-2, 1e-2, "tc vs simt")


def test_tc_multi_operand_strided_weights_adds_fp32_out():
    e = E()
    B, H, W, C, N = 2, 8, 16, 24, 40
    L0, a0 = mk_lazy((B, H, W, C), torch.bfloat16, 2, 3)
    L1, a1 = mk_lazy((B, H, W, C), torch.bfloat16, 0, 4)
    g = torch.Generator().manual_seed(5)
    w = (torch.randn(N, 2 * C, generator=g) / C ** 0.5).to(DEV)
    add1 = torch.randn(B, H // 2, W // 2, N, generator=g).to(DEV)
    add2 = torch.randn(B, H // 4, W // 4, N, generator=g).to(DEV)
    y = e.conv([e.Op(L0, C, e.WV(w, 0, 2 * C, 2)), e.Op(L1, C, e.WV(w, 1, 2 * C, 2))], N, (B, H, W),
               adds=[(add1, 1), (add2, 2)], out_dtype=e.F32)
    ref = bf(a0) @ bf(w[:, 0::2]).t() + bf(a1) @ bf(w[:, 1::2]).t()
    ref = ref + add1.repeat_interleave(2, 1).repeat_interleave(2, 2) + add2.repeat_interleave(4, 1).repeat_interleave(4, 2)
    assert y.dtype == torch.float32
    close(y, ref, RT, AT, "tc multi")


def test_tc_column_slices_and_sliced_output():
    e = E()
    B, H, W, C, N = 1, 4, 8, 16, 8
    L, a = mk_lazy((B, H, W, 2 * C), torch.bfloat16, 0, 6)
    w = torch.randn(N, 5 * C, generator=torch.Generator().manual_seed(7)).to(DEV) / 4
    out = torch.zeros(B, H, W, 3 * N, device=DEV, dtype=torch.float32)
    e.conv([e.Op(L, C, e.WV(w, 1, 5 * C, 5), 0), e.Op(L, C, e.WV(w, 3, 5 * C, 5), C)], N, (B, H, W), out=out, out_coff=N)
    ref = bf(a[..., :C]) @ bf(w[:, 1::5]).t() + bf(a[..., C:]) @ bf(w[:, 3::5]).t()
    close(out[..., N:2 * N], ref, RT, AT, "tc sliced out")
    assert float(out[..., :N].abs().max()) == 0 and float(out[..., 2 * N:].abs().max()) == 0


@pytest.mark.parametrize("C", [8, 32, 72])
def test_tc_dense3x3_nine_shifted_operands(C):
    e = E()
    B, H, W = 2, 9, 12
    L, a = mk_lazy((B, H, W, C), torch.bfloat16, 0, 8)
    g = torch.Generator().manual_seed(9)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    b = torch.randn(C, generator=g).to(DEV)
    ops = [e.Op(L, C, e.WV(w, ky * 3 + kx, C * 9, 9), 0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    y = e.conv(ops, C, (B, H, W), bias=b)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(bf(a).permute(0, 3, 1, 2), bf(w), b, padding=1).permute(0, 2, 3, 1)
    close(y.float(), ref, RT, AT, "tc conv3x3")


@pytest.mark.parametrize("shape", [(2, 9, 12), (3, 16, 28), (1, 1, 7), (2, 5, 2), (1, 40, 40), (2, 23, 3), (16, 56, 56)])
@pytest.mark.parametrize("C,N,act", [(32, 32, 2), (64, 64, 2), (16, 48, 1), (8, 8, 0), (24, 64, 2)])
def test_tc_dense3x3_slab_mode(shape, C, N, act):
    """the halo-slab mode of accx_pw_fwd_tc (one landed slab per filter row, three TMEM accumulators, column masks in
    the epilogue) against F.conv2d(padding=1) of the activated bf16 input: forward taps with statistics, and the
    transposed taps of the input gradient; multi-tile maps, several tiles per CTA (16x56x56: 392 tiles), maps narrower than the halo, P not a multiple of 128.
    Also against the nine-shifted-operand path it replaces (knob 18 = 2)."""
    from accx import _lib
    e = E()
    B, H, W = shape
    L, a = mk_lazy((B, H, W, C), torch.bfloat16, act, 21)
    g = torch.Generator().manual_seed(22)
    w = (torch.randn(N, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    ops = [e.Op(L, C, e.WV(w, ky * 3 + kx, C * 9, 9), 0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    stats = torch.zeros(2 * N, device=DEV)
    y = e.conv(ops, N, (B, H, W), stats=stats)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(bf(a).permute(0, 3, 1, 2), bf(w), None, padding=1).permute(0, 2, 3, 1)
    close(y.float(), ref, RT, AT, "slab conv3x3")
    yb = y.float()
    close(stats[:N], yb.sum((0, 1, 2)), 1e-3, 1e-3, "slab conv3x3 sum")
    close(stats[N:], (yb * yb).sum((0, 1, 2)), 1e-3, 1e-3, "slab conv3x3 sum of squares")
    _lib.call("accx_set_knob", 18, 2)
    try:
        y9 = e.conv(ops, N, (B, H, W))
    finally:
        _lib.call("accx_set_knob", 18, 0)
    close(y.float(), y9.float(), 1e-2, 1e-2, "slab vs nine operands")
    if C == N:           # transposed taps (input gradient of the same conv), fp32 output
        Ld, d = mk_lazy((B, H, W, N), torch.bfloat16, 0, 23)
        opsT = [e.Op(Ld, N, e.WV(w, ky * 3 + kx, 9, C * 9), 0, 1 - ky, 1 - kx) for ky in range(3) for kx in range(3)]
        dx = e.conv(opsT, C, (B, H, W), out_dtype=e.F32)
        refT = F.conv_transpose2d(bf(d).permute(0, 3, 1, 2), bf(w), None, padding=1).permute(0, 2, 3, 1)
        close(dx, refT, RT, AT, "slab conv3x3 transposed")


@pytest.mark.parametrize("K,N", [(32, 32), (96, 40), (64, 272), (128, 4352)])
@pytest.mark.parametrize("f32_out", [False, True])
def test_tc_residual_in_epilogue(K, N, f32_out):
    """accx_pw_fwd_tc_res: Y = A.W^T + R with R in the output dtype, added before the output is rounded; out-of-place
    and in place (R is Y); several column tiles (N = 272, 4352) and a ragged last chunk (N = 40)"""
    e = E()
    B, H, W = 2, 12, 20
    L, a = mk_lazy((B, H, W, K), torch.bfloat16, 2, 31)
    g = torch.Generator().manual_seed(32)
    w = (torch.randn(N, K, generator=g) / K ** 0.5).to(DEV)
    odt = torch.float32 if f32_out else torch.bfloat16
    r = torch.randn(B, H, W, N, generator=g).to(DEV).to(odt)
    ops = [e.Op(L, K, e.WV(w, 0, K, 1))]
    l0 = e.LAUNCHES
    y = e.conv(ops, N, (B, H, W), out_dtype=e.F32 if f32_out else None, residual=r)
    assert e.LAUNCHES - l0 == 1, "the residual was not fused into the contraction"
    ref = bf(a) @ bf(w).t() + r.float()
    close(y.float(), ref, RT, AT, "residual")
    y2 = r.clone()
    e.conv(ops, N, (B, H, W), out=y2, residual=y2)
    close(y2.float(), ref, RT, AT, "residual in place")
    assert torch.equal(y2, y)


@pytest.mark.parametrize("P_shape", [(2, 12, 20), (3, 7, 11), (4, 32, 32)])
@pytest.mark.parametrize("K,N,act", [(32, 96, 0), (96, 32, 2), (480, 64, 1), (128, 136, 2), (16, 256, 2), (8, 8, 0),
                                     (1088, 128, 2), (24, 40, 2), (128, 4352, 0)])
def test_tc_fp32_storage_split(P_shape, K, N, act):
    """fp32 storage on the bf16 tensor cores (three-term split, MODE 2 of pw_fwd_tc_kernel): the reference's own
    arithmetic is fp32, so the bound is the fp32 parity bound -- rtol 1e-3, atol 1e-5 * max -- against an fp64 product
    of the same fp32 inputs; statistics from the fp32 output"""
    e = E()
    B, H, W = P_shape
    L, a = mk_lazy((B, H, W, K), torch.float32, act, 41)
    w = (torch.randn(N, K, generator=torch.Generator().manual_seed(42)) / K ** 0.5).to(DEV)
    stats = torch.zeros(2 * N, device=DEV)
    from accx import _lib
    names = []
    orig = _lib.call
    _lib.call = lambda n, *a_: (names.append(n), orig(n, *a_))[1]
    try:
        y = e.conv([e.Op(L, K, e.WV(w, 0, K, 1))], N, (B, H, W), stats=stats)
    finally:
        _lib.call = orig
    assert names == ["accx_pw_fwd_tc_res"], names
    assert y.dtype == torch.float32
    ref = (a.double() @ w.double().t())
    close(y, ref, 1e-3, 1e-5, "fp32 split")
    assert float((y.double() - ref).abs().max() / ref.abs().max()) < 1e-5       # 3 x TF32: ~2^-21 per product
    close(stats[:N], ref.sum((0, 1, 2)), 1e-3, 1e-4, "fp32 split sum")
    close(stats[N:], (ref * ref).sum((0, 1, 2)), 1e-3, 1e-4, "fp32 split sum of squares")


def test_tc_fp32_storage_split_dense3x3_and_residual():
    """nine shifted fp32 operands (ResPath in fp32 storage), bias, residual in the epilogue"""
    e = E()
    B, H, W, C = 2, 9, 12, 32
    L, a = mk_lazy((B, H, W, C), torch.float32, 2, 43)
    g = torch.Generator().manual_seed(44)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    b = torch.randn(C, generator=g).to(DEV)
    r = torch.randn(B, H, W, C, generator=g).to(DEV)
    ops = [e.Op(L, C, e.WV(w, ky * 3 + kx, C * 9, 9), 0, ky - 1, kx - 1) for ky in range(3) for kx in range(3)]
    y = e.conv(ops, C, (B, H, W), bias=b, residual=r)
    torch.backends.cudnn.allow_tf32 = False
    ref = F.conv2d(a.double().permute(0, 3, 1, 2), w.double(), b.double(), padding=1).permute(0, 2, 3, 1) + r.double()
    close(y, ref, 1e-3, 1e-5, "fp32 split conv3x3 + residual")


def test_tc_large_tile_count_and_k_pipeline():
    """many M tiles, 17 N tiles (N=4352, the cnv72.conv1 shape) and a deep K loop (cnv72.hnc main: K=4352)"""
    e = E()
    B, H, W = 2, 28, 28
    L, a = mk_lazy((B, H, W, 128), torch.bfloat16, 0, 11)
    w = (torch.randn(4352, 128, generator=torch.Generator().manual_seed(12)) / 11.3).to(DEV)
    y = e.conv([e.Op(L, 128, e.WV(w, 0, 128, 1))], 4352, (B, H, W))
    close(y.float(), bf(a) @ bf(w).t(), RT, AT, "tc N=4352")
    L2, a2 = mk_lazy((B, H, W, 4352), torch.bfloat16, 2, 13)
    w2 = (torch.randn(128, 4352, generator=torch.Generator().manual_seed(14)) / 66.0).to(DEV)
    stats = torch.zeros(256, device=DEV)
    y2 = e.conv([e.Op(L2, 4352, e.WV(w2, 0, 4352, 1))], 128, (B, H, W), stats=stats)
    ref2 = bf(a2) @ bf(w2).t()
    close(y2.float(), ref2, RT, AT, "tc K=4352")
    close(stats[:128], ref2.sum((0, 1, 2)), 2e-2, 2e-2, "tc K=4352 stats")


@pytest.mark.parametrize("P_shape", [(2, 12, 20), (4, 32, 32), (1, 9, 15)])
@pytest.mark.parametrize("K,N,act", [(32, 96, 0), (96, 32, 2), (64, 64, 1), (128, 272, 2), (320, 128, 2), (16, 8, 0), (1088, 128, 2)])
def test_tc_wgrad(P_shape, K, N, act):
    e = E()
    B, H, W = P_shape
    L, a = mk_lazy((B, H, W, K), torch.bfloat16, act, 10)
    g = torch.Generator().manual_seed(11)
    dy = torch.randn(B, H, W, N, generator=g).to(DEV).to(torch.bfloat16)
    w = torch.zeros(N, K, device=DEV)
    gw = torch.zeros(N, K, device=DEV)
    n0 = e.LAUNCHES
    e.wgrad(e.Op(L, K, e.WV(w, 0, K, 1)), dy, N, (B, H, W), gw)
    ref = torch.einsum("bhwn,bhwk->nk", dy.float(), bf(a))
    close(gw, ref, RT, AT, "tc wgrad")


@pytest.mark.parametrize("K,N,act", [(32, 32, 2), (64, 64, 2), (96, 32, 1), (32, 96, 0), (64, 128, 2)])
def test_tc_wgrad_256_pixel_stages(K, N, act):
    """narrow weight gradients over many pixels run 256-pixel pipeline stages (P >= 1e5, at most three 64-channel blocks per
    stage); P not a multiple of 256; same result as the 128-pixel stages (knob 22) up to the order of the fp32 sums"""
    from accx import _lib
    e = E()
    B, H, W = 3, 181, 187               # 101541 pixels: 396 stages + a ragged tail
    L, a = mk_lazy((B, H, W, K), torch.bfloat16, act, 51)
    dy = torch.randn(B, H, W, N, generator=torch.Generator().manual_seed(52)).to(DEV).to(torch.bfloat16)
    w = torch.zeros(N, K, device=DEV)
    gw = torch.zeros(N, K, device=DEV)
    e.wgrad(e.Op(L, K, e.WV(w, 0, K, 1)), dy, N, (B, H, W), gw)
    ref = torch.einsum("bhwn,bhwk->nk", dy.double(), bf(a).double())
    close(gw, ref, 1e-3, 1e-4, "wgrad 256-pixel stages")
    gw2 = torch.zeros(N, K, device=DEV)
    _lib.call("accx_set_knob", 22, 128)
    try:
        e.wgrad(e.Op(L, K, e.WV(w, 0, K, 1)), dy, N, (B, H, W), gw2)
    finally:
        _lib.call("accx_set_knob", 22, 0)
    close(gw, gw2, 1e-4, 1e-5, "256- vs 128-pixel stages")


@pytest.mark.parametrize("ops_k,N,act,extra", [((32,), 32, 2, "res"), ((64,), 64, 2, "stats"), ((96,), 32, 1, "bias"),
                                               ((192,), 64, 2, "stats"), ((32, 32), 32, 2, "stats"), ((32,), 64, 0, "shift"),
                                               ((16,), 24, 2, "f32out")])
def test_tc_many_tiles_per_cta(ops_k, N, act, extra):
    """narrow contractions over many tiles (794: five or six per persistent CTA, so every pipeline stage, both TMEM
    accumulators and both epilogue groups wrap around several times; ragged last tile) with statistics / bias / residual /
    a shifted operand / fp32 output, against the torch restatement; two runs agree bit for bit in the outputs"""
    e = E()
    B, H, W = 3, 181, 187               # 101541 pixels: 794 tiles, a ragged last tile
    g = torch.Generator().manual_seed(61)
    Ls, As, ws = [], [], []
    for i, K in enumerate(ops_k):
        L, a = mk_lazy((B, H, W, K), torch.bfloat16, act, 62 + i)
        Ls.append(L), As.append(a)
    Kt = sum(ops_k)
    w = (torch.randn(N, Kt, generator=g) / Kt ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV) if extra == "bias" else None
    res = torch.randn(B, H, W, N, generator=g).to(DEV).to(torch.bfloat16) if extra == "res" else None
    dy, dx = (1, -1) if extra == "shift" else (0, 0)
    ops, k0 = [], 0
    for L, K in zip(Ls, ops_k):
        ops.append(e.Op(L, K, e.WV(w, k0, Kt, 1), 0, dy, dx))
        k0 += K
    kw = dict(bias=bias, out_dtype=e.F32 if extra == "f32out" else None)

    def run():
        stats = torch.zeros(2 * N, device=DEV) if extra in ("stats", "res") else None
        out = res.clone() if res is not None else None
        y = e.conv(ops, N, (B, H, W), stats=stats, residual=out, out=out, **kw) if res is not None else \
            e.conv(ops, N, (B, H, W), stats=stats, **kw)
        return y, stats

    y1, st1 = run()
    y2, st2 = run()
    assert torch.equal(y1, y2), "outputs differ between two runs"
    ref = torch.zeros(B, H, W, N, device=DEV)
    k0 = 0
    for a, K in zip(As, ops_k):
        ab = bf(a)
        if dy or dx:
            ab = F.pad(ab, (0, 0, 1, 1, 1, 1))[:, 1 + dy:1 + dy + H, 1 + dx:1 + dx + W]
        ref = ref + ab @ bf(w[:, k0:k0 + K]).t()
        k0 += K
    if bias is not None:
        ref = ref + bias
    if res is not None:
        ref = ref + res.float()
    close(y2.float(), ref, RT, AT, "many tiles per CTA")
    if st2 is not None:
        close(st2, st1, 1e-4, 1e-4, "statistics of two runs")
        close(st2[:N], y2.float().sum((0, 1, 2)), 2e-2, 2e-2, "stats sum")


@pytest.mark.parametrize("shape", [(3, 180, 188), (2, 12, 20), (1, 6, 6)])
@pytest.mark.parametrize("ops_k,N,act,extra", [((32,), 32, 2, "res"), ((32,), 32, 1, "stats"), ((32, 32), 32, 2, "stats"),
                                               ((32,), 64, 0, "bias"), ((96,), 32, 2, "adds"), ((64,), 32, 2, "stats"),
                                               ((16,), 48, 2, "stats"), ((32, 64), 64, 2, "adds"), ((8,), 16, 0, "res"),
                                               ((32,), 96, 0, "stats"), ((32,), 128, 2, "bias")])
def test_tc_pixel_folding(shape, ops_k, N, act, extra):
    """narrow contiguous contractions read two pixels per row ([P/2, 2C] views, block-diagonal weights; accx_pw_fwd_tc_res):
    outputs equal the unfolded launch (knob 23 = 1) bit for bit -- the extra products are exact zeros -- statistics agree up
    to the order of the fp32 sums, and both match the torch restatement; with bias, residual, several operands, a
    BatchNorm + LeakyReLU transform and nearest-upsampled addends (factor 2 and 4)"""
    from accx import _lib
    e = E()
    B, H, W = shape
    if extra == "adds" and (H % 4 or W % 4):
        pytest.skip("addends need a map divisible by 4")
    g = torch.Generator().manual_seed(71)
    Ls, As = [], []
    for i, K in enumerate(ops_k):
        L, a = mk_lazy((B, H, W, K), torch.bfloat16, act, 72 + i)
        Ls.append(L), As.append(a)
    Kt = sum(ops_k)
    w = (torch.randn(N, Kt, generator=g) / Kt ** 0.5).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV) if extra == "bias" else None
    res = torch.randn(B, H, W, N, generator=g).to(DEV).to(torch.bfloat16) if extra == "res" else None
    adds = [(torch.randn(B, H >> l, W >> l, N, generator=g).to(DEV), l) for l in (1, 2)] if extra == "adds" else []
    ops, k0 = [], 0
    for L, K in zip(Ls, ops_k):
        ops.append(e.Op(L, K, e.WV(w, k0, Kt, 1)))
        k0 += K

    def run():
        stats = torch.zeros(2 * N, device=DEV) if extra in ("stats", "res") else None
        if res is not None:
            out = res.clone()
            return e.conv(ops, N, (B, H, W), stats=stats, residual=out, out=out), stats
        return e.conv(ops, N, (B, H, W), bias=bias, adds=adds, stats=stats), stats

    y2, st2 = run()
    _lib.call("accx_set_knob", 23, 1)
    try:
        y1, st1 = run()
    finally:
        _lib.call("accx_set_knob", 23, 0)
    assert torch.equal(y1, y2), "pixel folding changed the outputs"
    ref = torch.zeros(B, H, W, N, device=DEV)
    k0 = 0
    for a, K in zip(As, ops_k):
        ref = ref + bf(a) @ bf(w[:, k0:k0 + K]).t()
        k0 += K
    if bias is not None:
        ref = ref + bias
    for t, l in adds:
        ref = ref + t.repeat_interleave(1 << l, 1).repeat_interleave(1 << l, 2)
    if res is not None:
        ref = ref + res.float()
    close(y2.float(), ref, RT, AT, "pixel folding")
    if st2 is not None:
        close(st2, st1, 1e-4, 1e-3, "statistics, folded vs plain rows")
        close(st2[:N], y2.float().sum((0, 1, 2)), 2e-2, 2e-2, "stats sum")
        close(st2[N:], y2.float().square().sum((0, 1, 2)), 2e-2, 2e-2, "stats sumsq")


@pytest.mark.parametrize("shape", [(3, 180, 188), (2, 12, 20), (1, 4, 4)])
@pytest.mark.parametrize("ops_k,N,act,levels", [((64,), 64, 2, (1,)), ((64,), 64, 2, (1, 2)), ((32,), 32, 2, (1,)), ((192,), 64, 1, (2, 1)),
                                               ((96,), 32, 2, (1, 2)), ((64, 64), 48, 2, (1,)), ((128,), 16, 0, (2,))])
def test_tc_first_addend_staged_through_shared_memory(shape, ops_k, N, act, levels):
    """narrow column tiles (<= 64 channels, N % 16 == 0) copy the first nearest-upsampled addend into shared memory with
    cp.async under the wait for the accumulator (knob 25); same outputs as direct loads bit for bit, with one and two
    addends, folded and plain rows, and against the torch restatement"""
    from accx import _lib
    e = E()
    B, H, W = shape
    g = torch.Generator().manual_seed(91)
    Ls, As = [], []
    for i, K in enumerate(ops_k):
        L, a = mk_lazy((B, H, W, K), torch.bfloat16, act, 92 + i)
        Ls.append(L), As.append(a)
    Kt = sum(ops_k)
    w = (torch.randn(N, Kt, generator=g) / Kt ** 0.5).to(DEV)
    adds = [(torch.randn(B, H >> l, W >> l, N, generator=g).to(DEV), l) for l in levels]
    ops, k0 = [], 0
    for L, K in zip(Ls, ops_k):
        ops.append(e.Op(L, K, e.WV(w, k0, Kt, 1)))
        k0 += K

    def run():
        stats = torch.zeros(2 * N, device=DEV)
        return e.conv(ops, N, (B, H, W), adds=adds, stats=stats), stats

    y2, st2 = run()
    _lib.call("accx_set_knob", 25, 1)
    try:
        y1, st1 = run()
    finally:
        _lib.call("accx_set_knob", 25, 0)
    assert torch.equal(y1, y2), "staging the addend changed the outputs"
    ref = torch.zeros(B, H, W, N, device=DEV)
    k0 = 0
    for a, K in zip(As, ops_k):
        ref = ref + bf(a) @ bf(w[:, k0:k0 + K]).t()
        k0 += K
    for t, l in adds:
        ref = ref + t.repeat_interleave(1 << l, 1).repeat_interleave(1 << l, 2)
    close(y2.float(), ref, RT, AT, "staged addend")
    close(st2, st1, 1e-4, 1e-3, "statistics, staged vs direct addend")


@pytest.mark.parametrize("shape", [(6, 180, 187), (2, 12, 20), (1, 6, 6)])
@pytest.mark.parametrize("K,N,act", [(32, 32, 2), (32, 64, 0), (96, 32, 2), (64, 32, 1), (16, 48, 2), (128, 32, 2), (32, 96, 2)])
def test_tc_wgrad_pixel_folding(shape, K, N, act):
    """weight gradients of narrow contiguous operands read two pixels per row and keep the two diagonal blocks of the
    [2N, 2K] product: same result as plain rows (knob 23 = 1) up to the order of the fp32 sums, and both match the fp64
    restatement; (32, 96) is not folded (N > 64)"""
    from accx import _lib
    e = E()
    B, H, W = shape
    L, a = mk_lazy((B, H, W, K), torch.bfloat16, act, 81)
    dy = torch.randn(B, H, W, N, generator=torch.Generator().manual_seed(82)).to(DEV).to(torch.bfloat16)
    w = torch.zeros(N, K, device=DEV)
    ref = torch.einsum("bhwn,bhwk->nk", dy.double(), bf(a).double())
    gw = torch.zeros(N, K, device=DEV)
    e.wgrad(e.Op(L, K, e.WV(w, 0, K, 1)), dy, N, (B, H, W), gw)
    torch.cuda.synchronize()
    gw1 = torch.zeros(N, K, device=DEV)
    _lib.call("accx_set_knob", 23, 1)
    try:
        e.wgrad(e.Op(L, K, e.WV(w, 0, K, 1)), dy, N, (B, H, W), gw1)
        torch.cuda.synchronize()
    finally:
        _lib.call("accx_set_knob", 23, 0)
    close(gw, ref, 1e-3, 1e-4, "folded weight gradient")
    close(gw, gw1, 1e-4, 1e-5, "folded vs plain rows")


def test_tc_wgrad_strided_weight_view_and_column_slices():
    """HANC layout: dW[n, e*J + j] for the max half (columns E..2E) of the pooled buffer"""
    e = E()
    B, H, W, Ein, N, J = 2, 8, 16, 24, 16, 5
    L, a = mk_lazy((B, H, W, 2 * Ein), torch.bfloat16, 0, 12)
    dy = torch.randn(B, H, W, N, generator=torch.Generator().manual_seed(13)).to(DEV).to(torch.bfloat16)
    w = torch.zeros(N, J * Ein, device=DEV)
    gw = torch.zeros(N, J * Ein, device=DEV)
    e.wgrad(e.Op(L, Ein, e.WV(w, 3, J * Ein, J), Ein), dy, N, (B, H, W), gw)
    ref = torch.einsum("bhwn,bhwk->nk", dy.float(), bf(a[..., Ein:]))
    close(gw[:, 3::J], ref, RT, AT, "tc wgrad strided")
    mask = torch.ones(J * Ein, dtype=torch.bool)
    mask[3::J] = False
    assert float(gw[:, mask.to(DEV)].abs().max()) == 0


@pytest.mark.parametrize("C", [8, 32, 72])
def test_tc_wgrad_dense3x3_taps(C):
    e = E()
    B, H, W = 2, 9, 12
    L, a = mk_lazy((B, H, W, C), torch.bfloat16, 0, 8)
    g = torch.Generator().manual_seed(9)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    dy = torch.randn(B, H, W, C, generator=g).to(DEV).to(torch.bfloat16)
    gw = torch.zeros_like(w)
    for ky in range(3):
        for kx in range(3):
            e.wgrad(e.Op(L, C, e.WV(w, ky * 3 + kx, C * 9, 9), 0, ky - 1, kx - 1), dy, C, (B, H, W), gw)
    torch.backends.cudnn.allow_tf32 = False
    w_ = w.clone().requires_grad_(True)
    F.conv2d(bf(a).permute(0, 3, 1, 2), w_, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    close(gw, w_.grad, RT, AT, "tc conv3x3 wgrad")


@pytest.mark.parametrize("C", [8, 32, 64, 72, 136])
def test_tc_wgrad_conv3x3_grouped_taps(C):
    """accx_pw_wgrad_taps_tc (several taps per pass over dY / the activation) against autograd of F.conv2d"""
    e = E()
    B, H, W = 2, 20, 24
    L, a = mk_lazy((B, H, W, C), torch.bfloat16, 2, 18)
    g = torch.Generator().manual_seed(19)
    w = (torch.randn(C, C, 3, 3, generator=g) / (3 * C ** 0.5)).to(DEV)
    dy = torch.randn(B, H, W, C, generator=g).to(DEV).to(torch.bfloat16)
    gw = torch.zeros_like(w)
    n0 = e.LAUNCHES
    e.wgrad_conv3x3(L, C, w, dy, C, (B, H, W), gw)
    assert e.LAUNCHES - n0 <= 3, "taps were not grouped"
    torch.backends.cudnn.allow_tf32 = False
    w_ = w.clone().requires_grad_(True)
    F.conv2d(bf(a).permute(0, 3, 1, 2), w_, None, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    close(gw, w_.grad, RT, AT, "tc conv3x3 grouped wgrad")
