"""Drop-in replacements for the five hot-path modules of the reference
(/root/reference/ACC_UNet/ACC_UNet.py): same class names, constructor signatures, submodule
tree (hence identical state_dict keys and default-initialiser RNG order), forward signatures
and train()/eval() behaviour -- but `forward` runs the accx sm_100a kernels through the C ABI
(include/accx.h) and `backward` is hand-written on the same kernels.

    ChannelSELayer(num_channels)                      ACC_UNet.py:9-49
    HANCLayer(in_chnl, out_chnl, k)                   ACC_UNet.py:53-142
    Conv2d_batchnorm(in, out, kernel_size, ...)       ACC_UNet.py:146-186   (1x1 only, as used by MLFC)
    HANCBlock(n_filts, out_channels, k=3, inv_fctr=3) ACC_UNet.py:224-286
    ResPath(in_chnls, n_lvl)                          ACC_UNet.py:290-328
    MLFC(f1, f2, f3, f4, lenn=1)                      ACC_UNet.py:332-527  (+ `variant` = "w" | "lite":
                                                      ACC_UNet_w.py:354,497-522; ACC_UNet_lite.py:424-427)

Inputs/outputs are NCHW-shaped tensors like the reference's (channels_last strides are
consumed and produced without a copy); fp32 or bf16 activations, fp32 parameters.
There is no CPU path: a non-CUDA input raises.
"""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from . import engine as E
from .engine import Arena, Lazy, Op, WV


# =========================================================================================
# autograd glue shared by all modules
# =========================================================================================
class _ModuleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mod, n_in, *tensors):
        xs = tensors[:n_in]
        E.require_cuda(xs[0])
        need = any(ctx.needs_input_grad[2:])
        xs_n = [E.to_nhwc(x.detach()) for x in xs]
        outs, saved = mod._fwd(xs_n, mod.training, need)
        ctx.mod, ctx.saved, ctx.n_in = mod, saved, n_in
        ctx.params = tensors[n_in:]
        ctx.training = mod.training
        ctx.in_need = ctx.needs_input_grad[2:2 + n_in]
        res = tuple(E.to_nchw_view(o) for o in outs)
        return res if len(res) > 1 else res[0]

    @staticmethod
    def backward(ctx, *douts):
        saved, mod = ctx.saved, ctx.mod
        ctx.saved = None
        if E.BWD_START_HOOK is not None:
            E.BWD_START_HOOK(mod)
        ref = saved["out_like"]
        dn = []
        for d, like in zip(douts, ref):
            if d is None:
                d = torch.zeros(like[0], dtype=like[1], device=like[2]).permute(0, 3, 1, 2)
            if d.dtype != like[1]:
                d = d.to(like[1])
            dn.append(E.to_nhwc(d))
        E.BWD_DEPTH[0] += 1
        try:
            dxs, grads = mod._bwd(saved, dn, ctx.in_need)
        finally:
            E.BWD_DEPTH[0] -= 1
            E.module_backward_end()      # weight gradients run on a side stream (engine.side_stream)
        gp = E.param_grads(ctx.params, grads)
        gx = [None if dx is None else E.to_nchw_view(dx) for dx in dxs]
        return (None, None, *gx, *gp)


class _GroupFn(torch.autograd.Function):
    """Independent CHAINS of single-input / single-output accx modules (the model's four ResPaths, the bottleneck
    pair cnv51 -> cnv52) as ONE autograd node: the chains run forward and backward on parallel stream lanes."""

    @staticmethod
    def forward(ctx, chains, *tensors):
        n = len(chains)
        xs = tensors[:n]
        E.require_cuda(xs[0])
        need = any(ctx.needs_input_grad[1:])
        outs, saved = [None] * n, [None] * n
        xs_n = [E.to_nhwc(x.detach()) for x in xs]
        with E.fork_lanes(n) as lanes:
            for i, chain in enumerate(chains):
                with lanes.lane(i):
                    cur, sv = xs_n[i], []
                    for m in chain:
                        o, s_ = m._fwd([cur], m.training, need)
                        cur = o[0]
                        sv.append(s_)
                    outs[i], saved[i] = cur, sv
        ctx.chains, ctx.saved = chains, saved
        ctx.params = tensors[n:]
        ctx.training = all(m.training for c in chains for m in c)
        ctx.in_need = ctx.needs_input_grad[1:1 + n]
        return tuple(E.to_nchw_view(o) for o in outs)

    @staticmethod
    def backward(ctx, *douts):
        chains, saved, n = ctx.chains, ctx.saved, len(ctx.chains)
        ctx.saved = None
        if E.BWD_START_HOOK is not None:
            E.BWD_START_HOOK(chains)
        dn = []
        for d, sv in zip(douts, saved):
            like = sv[-1]["out_like"][0]
            if d is None:
                d = torch.zeros(like[0], dtype=like[1], device=like[2]).permute(0, 3, 1, 2)
            if d.dtype != like[1]:
                d = d.to(like[1])
            dn.append(E.to_nhwc(d))
        # gradient accumulators are zeroed on the caller's stream, before the fork
        pools = [[E.GradPool(m.parameters()) for m in chain] for chain in chains]
        gx = [None] * n
        E.BWD_DEPTH[0] += 1
        try:
            with E.fork_lanes(n) as lanes:
                for i, chain in enumerate(chains):
                    with lanes.lane(i):
                        cur = dn[i]
                        for j in reversed(range(len(chain))):
                            need_in = True if j > 0 else ctx.in_need[i]
                            dxs, _ = chain[j]._bwd(saved[i][j], [cur], [need_in], grads=pools[i][j])
                            cur = dxs[0]
                        gx[i] = cur
        finally:
            E.BWD_DEPTH[0] -= 1
            E.module_backward_end()
        grads = {}
        for chain_pools in pools:
            for gp in chain_pools:
                grads.update(gp)
        gp_all = E.param_grads(ctx.params, grads)
        return (None, *[None if g is None else E.to_nchw_view(g) for g in gx], *gp_all)


def run_parallel(chains, xs):
    """y_i = chain_i(xs[i]) for independent chains (lists) of accx modules, issued concurrently (see _GroupFn)"""
    chains = [list(c) if isinstance(c, (list, tuple)) else [c] for c in chains]
    params = [p for c in chains for m in c for p in m.parameters()]
    return _GroupFn.apply(chains, *xs, *params)


class _MaxPool2Fn(torch.autograd.Function):
    """nn.MaxPool2d(2) between the encoder levels (ACC_UNet.py:552,608-618) on the accx kernels; backward
    recomputes the arg-max from the saved input (no index tensor)."""

    @staticmethod
    def forward(ctx, x):
        E.require_cuda(x)
        xn = E.to_nhwc(x.detach())
        if xn.shape[1] % 2 or xn.shape[2] % 2:
            raise ValueError(f"accx MaxPool2d(2) needs even H, W, got {xn.shape[1]}x{xn.shape[2]}")
        ctx.xn = xn
        return E.to_nchw_view(E.maxpool2(xn))

    @staticmethod
    def backward(ctx, dy):
        xn, ctx.xn = ctx.xn, None
        dn = E.to_nhwc(dy if dy.dtype == xn.dtype else dy.to(xn.dtype))
        return E.to_nchw_view(E.maxpool2_bwd(xn, dn))


def maxpool2(x):
    return _MaxPool2Fn.apply(x)


class _OutConvFn(torch.autograd.Function):
    """the model's final 1x1 conv (ACC_UNet.py:596-599,653) -> fp32 logits, on the tiny-channel contraction kernels"""

    @staticmethod
    def forward(ctx, x, weight, bias):
        E.require_cuda(x)
        xn = E.to_nhwc(x.detach())
        B, H, W, C = xn.shape
        N = weight.shape[0]
        w = E.f32(weight)
        y = E.conv([Op(Lazy(xn), C, WV(w, 0, C, 1))], N, (B, H, W), bias=E.f32(bias), out_dtype=E.F32)
        ctx.xn, ctx.weight, ctx.bias = xn, weight, bias
        return E.to_nchw_view(y)

    @staticmethod
    def backward(ctx, dy):
        xn, weight, bias = ctx.xn, ctx.weight, ctx.bias
        ctx.xn = None
        B, H, W, C = xn.shape
        N = weight.shape[0]
        dims = (B, H, W)
        w = E.f32(weight)
        dn = E.to_nhwc(dy.float())
        grads = E.GradPool([weight] + ([bias] if bias is not None else []))
        gw, gb = None, None
        if ctx.needs_input_grad[1]:
            gw = E.grad_buf(grads, weight)
            E.wgrad(Op(Lazy(xn), C, WV(w, 0, C, 1)), dn, N, dims, gw)
        if bias is not None and ctx.needs_input_grad[2]:
            gb = E.grad_buf(grads, bias)
            torch.sum(dn.view(-1, N), dim=0, out=gb)
        dx = None
        if ctx.needs_input_grad[0]:
            dl = dn if dn.dtype == xn.dtype else dn.to(xn.dtype)
            dx = E.to_nchw_view(E.conv([Op(Lazy(dl), N, WV(w, 0, 1, C))], C, dims))
        gw, gb = E.param_grads([weight, bias], grads)
        return dx, gw, gb


def out_conv(x, conv: nn.Conv2d):
    return _OutConvFn.apply(x, conv.weight, conv.bias)


class _UpCatFn(torch.autograd.Function):
    """torch.cat([ConvTranspose2d(Cin, Co, 2, stride 2)(x), skip], dim=1) of the decoder (ACC_UNet.py:578-599,620-631)
    on accx kernels: the transposed conv is one pointwise contraction [P, Cin] x [Cin, 4*Co] (weight [Cin, Co, 2, 2]
    through a strided view), interleaved into the left half of the concat buffer; the skip is copied into the right."""

    @staticmethod
    def forward(ctx, x, skip, weight, bias):
        E.require_cuda(x)
        xn, sn = E.to_nhwc(x.detach()), E.to_nhwc(skip.detach())
        B, H, W, Cin = xn.shape
        Co, Cs = weight.shape[1], sn.shape[3]
        if tuple(weight.shape) != (Cin, Co, 2, 2) or tuple(sn.shape[:3]) != (B, 2 * H, 2 * W) or Co % 2 or sn.dtype != xn.dtype:
            raise ValueError(f"accx up+cat: x {tuple(x.shape)}, skip {tuple(skip.shape)}, weight {tuple(weight.shape)}")
        w = E.f32(weight)
        temp = E.conv([Op(Lazy(xn), Cin, WV(w, 0, 1, 4 * Co))], 4 * Co, (B, H, W))
        out = torch.empty((B, 2 * H, 2 * W, Co + Cs), dtype=xn.dtype, device=xn.device)
        E.upshuffle(temp, E.f32(bias), out, Co, forward=True)
        E.copy_cols(sn, 0, out, Co, Cs)
        ctx.xn, ctx.weight, ctx.bias, ctx.dims = xn, weight, bias, (B, H, W, Cin, Co, Cs)
        return E.to_nchw_view(out)

    @staticmethod
    def backward(ctx, dout):
        xn, weight, bias = ctx.xn, ctx.weight, ctx.bias
        ctx.xn = None
        B, H, W, Cin, Co, Cs = ctx.dims
        dn = E.to_nhwc(dout if dout.dtype == xn.dtype else dout.to(xn.dtype))
        w = E.f32(weight)
        grads = E.GradPool([weight] + ([bias] if bias is not None else []))
        gw = gb = None
        if bias is not None and ctx.needs_input_grad[3]:
            gb = E.grad_buf(grads, bias)
        dtemp = torch.empty((B, H, W, 4 * Co), dtype=xn.dtype, device=xn.device)
        E.upshuffle(dtemp, None, dn, Co, forward=False, dbias=gb)         # + the bias gradient, in the same pass
        dskip = None
        if ctx.needs_input_grad[1]:
            dskip = torch.empty((B, 2 * H, 2 * W, Cs), dtype=xn.dtype, device=xn.device)
            E.copy_cols(dn, Co, dskip, 0, Cs)
            dskip = E.to_nchw_view(dskip)
        if ctx.needs_input_grad[2]:
            gw = E.grad_buf(grads, weight)
            # dW[ci, n'] = sum_p x[p, ci] * dtemp[p, n']: dtemp as the operand, x as "dY", so that the accumulated
            # rows are contiguous in the reference's [Cin, Co*4] layout (vector atomics)
            E.wgrad(Op(Lazy(dtemp), 4 * Co, WV(w, 0, 4 * Co, 1)), xn, Cin, (B, H, W), gw)
        dx = None
        if ctx.needs_input_grad[0]:
            dx = E.to_nchw_view(E.conv([Op(Lazy(dtemp), 4 * Co, WV(w, 0, 4 * Co, 1))], Cin, (B, H, W)))
        gw, gb = E.param_grads([weight, bias], grads)
        return dx, dskip, gw, gb


def up_cat(x, skip, up: nn.ConvTranspose2d):
    """cat([up(x), skip], dim=1) for up = ConvTranspose2d(Cin, Co, kernel_size=2, stride=2)"""
    if tuple(up.kernel_size) != (2, 2) or tuple(up.stride) != (2, 2) or tuple(up.padding) != (0, 0) or up.groups != 1:
        raise NotImplementedError("accx up_cat implements ConvTranspose2d(kernel_size=2, stride=2) (all the reference builds)")
    return _UpCatFn.apply(x, skip, up.weight, up.bias)


class _AccxModule(nn.Module):
    def _run(self, *xs):
        params = [p for p in self.parameters()]
        return _ModuleFn.apply(self, len(xs), *xs, *params)


def _out_like(outs):
    return [(tuple(o.shape), o.dtype, o.device) for o in outs]


def _w(p):
    return E.f32(p)


# =========================================================================================
# building blocks shared by several modules (forward returns what backward needs)
# =========================================================================================
def _pw_bn(ops: List[Op], N, dims, conv: nn.Conv2d, bn: nn.BatchNorm2d, act, ar: Arena, training, adds=()):
    st = ar.take(2 * N) if training else None
    # the conv bias is not added to y: a per-channel constant cancels in the BatchNorm that follows, so it is
    # folded into the BN affine / running mean by accx_bn_finalize and never touches the big tensor
    y = E.conv(ops, N, dims, adds=adds, stats=st)
    return E.bn_lazy(y, st, bn, act, ar, training, conv_bias=conv.bias)


def _hanc_core_fwd(hnc, L2: Lazy, ar: Arena, training):
    """HANCLayer on a lazy input.  Split form: pooled maps get their own low-resolution
    contraction, results are nearest-upsample-added in the main contraction's epilogue."""
    B, H, W, Ein = L2.y.shape
    k = hnc.k
    J = 2 * k - 1
    C = hnc.cnv.out_channels
    if H % (1 << (k - 1)) or W % (1 << (k - 1)):
        raise ValueError(f"HANCLayer(k={k}) needs H, W divisible by {1 << (k - 1)}, got {H}x{W}")
    wh = _w(hnc.cnv.weight)            # [C, J*Ein] with K index e*J + j
    if E.RECORD is not None:
        E.RECORD[("hanc", id(hnc))] = L2
    pools = E.hanc_pools(L2, k)
    # bf16 storage: coarsest level first; every level adds the (nearest-upsampled) sum of the coarser ones in its own
    # epilogue, so the full-resolution contraction reads ONE addend whatever k is (addend loads are a chain of L2 round
    # trips per chunk in the epilogue: K = N = 32 at 16x224x224 costs 32 us without addends and 100 us with three).
    # fp32 storage (the parity mode) keeps one addend per level: the association of the fp32 sums stays what the
    # reference-pinned tolerances were measured with.
    chain = L2.y.dtype == torch.bfloat16
    adds, r = [], None
    for l in range(k - 1, 0, -1):
        Pl = Lazy(pools[l - 1])
        r = E.conv([Op(Pl, Ein, WV(wh, l, J * Ein, J), 0), Op(Pl, Ein, WV(wh, k - 1 + l, J * Ein, J), Ein)],
                   C, (B, H >> l, W >> l), out_dtype=E.F32, adds=[(r, 1)] if (chain and r is not None) else ())
        adds.insert(0, (r, l))
    if chain:
        adds = adds[:1]
    L3 = _pw_bn([Op(L2, Ein, WV(wh, 0, J * Ein, J))], C, (B, H, W), hnc.cnv, hnc.bn, 2, ar, training, adds)
    return L3, pools


def _hanc_core_bwd(hnc, L2: Lazy, pools, dy3: torch.Tensor, grads, ar: Arena, need_da=True, fuse_bn=False):
    """dy3 = gradient w.r.t. the raw conv output.  Returns the gradient w.r.t. the activated input
    (fuse_bn: -> (gradient, BN-backward sums of L2's BatchNorm or None), see engine.hanc_unpool_bnred)."""
    B, H, W, Ein = L2.y.shape
    k = hnc.k
    J = 2 * k - 1
    C = hnc.cnv.out_channels
    wh = _w(hnc.cnv.weight)
    gw = E.grad_buf(grads, hnc.cnv.weight)
    if gw is not None:
        E.wgrad(Op(L2, Ein, WV(wh, 0, J * Ein, J)), dy3, C, (B, H, W), gw)
    da2 = None
    fused = fuse_bn and need_da and E.hanc_unpool_fusable(L2, k)
    dPs = []
    # the full-resolution input gradient and the (small) per-level chains only meet in the unpool step:
    # lane 0 = main contraction, lane l = level l (block sums of dY -> weight gradients -> dP contractions)
    with E.fork_lanes(k if fused else 1) as lanes:
        if need_da:
            with lanes.lane(0):
                da2 = E.conv([Op(Lazy(dy3), C, WV(wh, 0, J, J * Ein))], Ein, (B, H, W))
        for l in range(1, k):
            with lanes.lane(l if fused else 0):
                dims_l = (B, H >> l, W >> l)
                dR = E.pool_sum(dy3, l, 1.0)                         # block sums of dY at the pooled resolution
                Pl = Lazy(pools[l - 1])
                if gw is not None:
                    E.wgrad(Op(Pl, Ein, WV(wh, l, J * Ein, J), 0), dR, C, dims_l, gw)
                    E.wgrad(Op(Pl, Ein, WV(wh, k - 1 + l, J * Ein, J), Ein), dR, C, dims_l, gw)
                if need_da:
                    dP = torch.empty((B, H >> l, W >> l, 2 * Ein), dtype=torch.float32, device=dy3.device)   # fp32
                    E.conv([Op(Lazy(dR), C, WV(wh, l, J, J * Ein))], Ein, dims_l, out=dP, out_coff=0)
                    E.conv([Op(Lazy(dR), C, WV(wh, k - 1 + l, J, J * Ein))], Ein, dims_l, out=dP, out_coff=Ein)
                    if fused:
                        dPs.append(dP)
                    else:
                        E.hanc_unpool_bwd(L2, l, dP, da2, accumulate=True)
    if fuse_bn:
        return da2, (E.hanc_unpool_bnred(L2, dPs, da2, ar) if fused else None)
    return da2


# =========================================================================================
class ChannelSELayer(_AccxModule):
    """Squeeze-and-excitation gate followed by BatchNorm + LeakyReLU (ACC_UNet.py:9-49)."""

    def __init__(self, num_channels):
        super().__init__()
        self.gp_avg_pool = nn.AdaptiveAvgPool2d(1)
        self.reduction_ratio = 8
        num_channels_reduced = num_channels // self.reduction_ratio
        self.fc1 = nn.Linear(num_channels, num_channels_reduced, bias=True)
        self.fc2 = nn.Linear(num_channels_reduced, num_channels, bias=True)
        self.act = nn.LeakyReLU()
        self.sigmoid = nn.Sigmoid()
        self.bn = nn.BatchNorm2d(num_channels)

    def forward(self, inp):
        return self._run(inp)

    def _fwd(self, xs, training, need):
        ar = Arena(xs[0].device)
        out, c = E.se_fwd(Lazy(xs[0]), self, ar, training)
        return [out], ({"se": c, "out_like": _out_like([out])} if need else None)

    def _bwd(self, s, douts, in_need):
        grads = E.GradPool(self.parameters())
        ar = Arena(douts[0].device)
        da = E.se_bwd(s["se"], douts[0], grads, ar)
        return [da], grads


# =========================================================================================
class HANCLayer(_AccxModule):
    """Hierarchical aggregation of neighbourhood context + 1x1 conv + BN + LeakyReLU (ACC_UNet.py:53-142)."""

    def __init__(self, in_chnl, out_chnl, k):
        super().__init__()
        self.k = k
        self.cnv = nn.Conv2d((2 * k - 1) * in_chnl, out_chnl, kernel_size=(1, 1))
        self.act = nn.LeakyReLU()
        self.bn = nn.BatchNorm2d(out_chnl)

    def forward(self, inp):
        return self._run(inp)

    def _fwd(self, xs, training, need):
        ar = Arena(xs[0].device)
        L2 = Lazy(xs[0])
        L3, pools = _hanc_core_fwd(self, L2, ar, training)
        out = E.materialize(L3)
        return [out], ({"L2": L2, "L3": L3, "pools": pools, "out_like": _out_like([out])} if need else None)

    def _bwd(self, s, douts, in_need):
        grads = E.GradPool(self.parameters())
        ar = Arena(douts[0].device)
        dy3 = E.bn_bwd(s["L3"], douts[0], grads, ar, out=torch.empty_like(s["L3"].y), conv=self.cnv)
        da = _hanc_core_bwd(self, s["L2"], s["pools"], dy3, grads, ar, need_da=in_need[0])
        return [da], grads


# =========================================================================================
class Conv2d_batchnorm(_AccxModule):
    """conv -> BN -> LeakyReLU -> SE (ACC_UNet.py:146-186).  Only 1x1 kernels (all the reference builds)."""

    def __init__(self, num_in_filters, num_out_filters, kernel_size, stride=(1, 1), activation="LeakyReLU"):
        super().__init__()
        if tuple(kernel_size) != (1, 1) or tuple(stride) != (1, 1):
            raise NotImplementedError("accx Conv2d_batchnorm implements the 1x1 / stride-1 case used by MLFC")
        self.activation = nn.LeakyReLU()
        self.conv1 = nn.Conv2d(in_channels=num_in_filters, out_channels=num_out_filters, kernel_size=kernel_size,
                               stride=stride, padding="same")
        self.batchnorm = nn.BatchNorm2d(num_out_filters)
        self.sqe = ChannelSELayer(num_out_filters)

    def forward(self, x):
        return self._run(x)

    # helpers used by MLFC as well -----------------------------------------------------
    def _core_fwd(self, ops, dims, ar, training, adds=(), residual=None, mix=None, stats=None, mix_param=None):
        N = self.conv1.out_channels
        L = _pw_bn(ops, N, dims, self.conv1, self.batchnorm, 2, ar, training, adds)
        out, c = E.se_fwd(L, self.sqe, ar, training, residual=residual, mix=mix, stats=stats, mix_param=mix_param)
        return out, (L, c)

    def _core_bwd(self, saved, dout, grads, ar, gmix=None):
        """-> gradient w.r.t. the raw conv output"""
        L, c = saved
        da, sums = E.se_bwd(c, dout, grads, ar, bn_sums=True, gmix=gmix)
        return E.bn_bwd(L, da, grads, ar, sums=sums, conv=self.conv1)

    def _fwd(self, xs, training, need):
        x = xs[0]
        B, H, W, K = x.shape
        ar = Arena(x.device)
        w = _w(self.conv1.weight)
        X = Lazy(x)
        out, sv = self._core_fwd([Op(X, K, WV(w, 0, K, 1))], (B, H, W), ar, training)
        return [out], ({"X": X, "sv": sv, "out_like": _out_like([out])} if need else None)

    def _bwd(self, s, douts, in_need):
        grads = E.GradPool(self.parameters())
        ar = Arena(douts[0].device)
        X = s["X"]
        B, H, W, K = X.y.shape
        N = self.conv1.out_channels
        dy = self._core_bwd(s["sv"], douts[0], grads, ar)
        w = _w(self.conv1.weight)
        gw = E.grad_buf(grads, self.conv1.weight)
        if gw is not None:
            E.wgrad(Op(X, K, WV(w, 0, K, 1)), dy, N, (B, H, W), gw)
        dx = E.conv([Op(Lazy(dy), N, WV(w, 0, 1, K))], K, (B, H, W)) if in_need[0] else None
        return [dx], grads


# =========================================================================================
class HANCBlock(_AccxModule):
    """1x1 expand -> dw3x3 -> HANC -> +inp, BN -> 1x1 project -> SE (ACC_UNet.py:224-286)."""

    def __init__(self, n_filts, out_channels, k=3, inv_fctr=3):
        super().__init__()
        self.conv1 = nn.Conv2d(n_filts, n_filts * inv_fctr, kernel_size=1)
        self.norm1 = nn.BatchNorm2d(n_filts * inv_fctr)
        self.conv2 = nn.Conv2d(n_filts * inv_fctr, n_filts * inv_fctr, kernel_size=3, padding=1,
                               groups=n_filts * inv_fctr)
        self.norm2 = nn.BatchNorm2d(n_filts * inv_fctr)
        self.hnc = HANCLayer(n_filts * inv_fctr, n_filts, k)
        self.norm = nn.BatchNorm2d(n_filts)
        self.conv3 = nn.Conv2d(n_filts, out_channels, kernel_size=1)
        self.norm3 = nn.BatchNorm2d(out_channels)
        self.sqe = ChannelSELayer(out_channels)
        self.activation = nn.LeakyReLU()

    def forward(self, inp):
        return self._run(inp)

    def _fwd(self, xs, training, need):
        x = xs[0]
        B, H, W, C = x.shape
        dims = (B, H, W)
        Ex = self.conv1.out_channels
        Cout = self.conv3.out_channels
        ar = Arena(x.device)
        X = Lazy(x)
        w1 = _w(self.conv1.weight)
        L1 = _pw_bn([Op(X, C, WV(w1, 0, C, 1))], Ex, dims, self.conv1, self.norm1, 2, ar, training)
        st2 = ar.take(2 * Ex) if training else None
        y2 = E.dw_fwd(L1, _w(self.conv2.weight), _w(self.conv2.bias), st2)
        L2 = E.bn_lazy(y2, st2, self.norm2, 2, ar, training)
        L3, pools = _hanc_core_fwd(self.hnc, L2, ar, training)
        stz = ar.take(2 * C) if training else None
        z = E.add_fwd(L3, x, stz)
        L4 = E.bn_lazy(z, stz, self.norm, 1, ar, training)
        w3 = _w(self.conv3.weight)
        L5 = _pw_bn([Op(L4, C, WV(w3, 0, C, 1))], Cout, dims, self.conv3, self.norm3, 2, ar, training)
        out, sec = E.se_fwd(L5, self.sqe, ar, training)
        saved = None
        if need:
            saved = {"X": X, "L1": L1, "L2": L2, "L3": L3, "pools": pools, "L4": L4, "L5": L5, "se": sec,
                     "out_like": _out_like([out])}
        return [out], saved

    def _bwd(self, s, douts, in_need, grads=None):
        grads = E.GradPool(self.parameters()) if grads is None else grads
        dout = douts[0]
        ar = Arena(dout.device)
        X, L1, L2, L3, L4, L5 = s["X"], s["L1"], s["L2"], s["L3"], s["L4"], s["L5"]
        B, H, W, C = X.y.shape
        dims = (B, H, W)
        Ex, Cout = self.conv1.out_channels, self.conv3.out_channels
        w1, w3 = _w(self.conv1.weight), _w(self.conv3.weight)
        # SE -> norm3 -> conv3
        da5, sums5 = E.se_bwd(s["se"], dout, grads, ar, bn_sums=True)
        dy5 = E.bn_bwd(L5, da5, grads, ar, sums=sums5, conv=self.conv3)
        g3 = E.grad_buf(grads, self.conv3.weight)
        if g3 is not None:
            E.wgrad(Op(L4, C, WV(w3, 0, C, 1)), dy5, Cout, dims, g3)
        da4 = E.conv([Op(Lazy(dy5), Cout, WV(w3, 0, 1, C))], C, dims)
        # norm(x + inp): dz feeds both the HANC branch and the residual
        dz = E.bn_bwd(L4, da4, grads, ar)
        dy3 = E.bn_bwd(L3, dz, grads, ar, out=torch.empty_like(dz), conv=self.hnc.cnv)
        da2, sums2 = _hanc_core_bwd(self.hnc, L2, s["pools"], dy3, grads, ar, fuse_bn=True)
        dy2 = E.bn_bwd(L2, da2, grads, ar, sums=sums2, conv=self.conv2)
        # depthwise
        g2 = E.grad_buf(grads, self.conv2.weight)
        if g2 is not None:
            E.dw_wgrad(L1, dy2, g2)
        if E.dw_dgrad_bnred_ok(L1, dy2):       # input gradient + norm1's backward reduction in one pass
            da1, sums1 = E.dw_dgrad_bnred(L1, dy2, _w(self.conv2.weight), ar)
        else:
            da1, sums1 = E.dw_fwd(Lazy(dy2), _w(self.conv2.weight), None, None, flip=True), None
        dy1 = E.bn_bwd(L1, da1, grads, ar, sums=sums1, conv=self.conv1)
        # conv1
        g1 = E.grad_buf(grads, self.conv1.weight)
        if g1 is not None:
            E.wgrad(Op(X, C, WV(w1, 0, C, 1)), dy1, Ex, dims, g1)
        dx = None
        if in_need[0]:
            dx = E.conv([Op(Lazy(dy1), Ex, WV(w1, 0, 1, C))], C, dims, residual=dz)      # + the skip branch, in the epilogue
        return [dx], grads


# =========================================================================================
class ResPath(_AccxModule):
    """n_lvl x [x += SE(lrelu(BN(conv3x3(x))))], then BN(lrelu(BN(x))) (ACC_UNet.py:290-328).
    NB: the attribute named `sqe` is a BatchNorm2d, as in the reference."""

    def __init__(self, in_chnls, n_lvl):
        super().__init__()
        self.convs = nn.ModuleList([])
        self.bns = nn.ModuleList([])
        self.sqes = nn.ModuleList([])
        self.bn = nn.BatchNorm2d(in_chnls)
        self.act = nn.LeakyReLU()
        self.sqe = nn.BatchNorm2d(in_chnls)
        for i in range(n_lvl):
            self.convs.append(nn.Conv2d(in_chnls, in_chnls, kernel_size=(3, 3), padding=1))
            self.bns.append(nn.BatchNorm2d(in_chnls))
            self.sqes.append(ChannelSELayer(in_chnls))

    def forward(self, x):
        return self._run(x)

    @staticmethod
    def _taps(X: Lazy, w, C, transpose=False):
        ops = []
        for ky in range(3):
            for kx in range(3):
                t = ky * 3 + kx
                if not transpose:      # y[p] += W[:, :, ky, kx] . x[p + (ky-1, kx-1)]
                    ops.append(Op(X, C, WV(w, t, C * 9, 9), 0, ky - 1, kx - 1))
                else:                  # dx[q] += W[:, :, ky, kx]^T . dy[q - (ky-1, kx-1)]
                    ops.append(Op(X, C, WV(w, t, 9, C * 9), 0, 1 - ky, 1 - kx))
        return ops

    def _fwd(self, xs, training, need):
        x = xs[0]
        B, H, W, C = x.shape
        dims = (B, H, W)
        ar = Arena(x.device)
        n = len(self.convs)
        levels = []
        st_last = ar.take(2 * C) if training else None
        if n == 0 and training:
            E.materialize(Lazy(x), stats=st_last, stats_only=True)
        for i in range(n):
            X = Lazy(x)
            w = _w(self.convs[i].weight)
            L = _pw_bn(self._taps(X, w, C), C, dims, self.convs[i], self.bns[i], 2, ar, training)
            x, sec = E.se_fwd(L, self.sqes[i], ar, training, residual=x,
                              stats=st_last if (i == n - 1 and training) else None)
            levels.append((X, L, sec))
        La = E.bn_lazy(x, st_last, self.bn, 2, ar, training)
        stu = ar.take(2 * C) if training else None
        u = E.materialize(La, stats=stu)
        Lb = E.bn_lazy(u, stu, self.sqe, 1, ar, training)
        out = E.materialize(Lb)
        saved = {"levels": levels, "La": La, "Lb": Lb, "out_like": _out_like([out])} if need else None
        return [out], saved

    def _bwd(self, s, douts, in_need, grads=None):
        grads = E.GradPool(self.parameters()) if grads is None else grads
        ar = Arena(douts[0].device)
        La, Lb = s["La"], s["Lb"]
        B, H, W, C = La.y.shape
        dims = (B, H, W)
        du = E.bn_bwd(Lb, douts[0], grads, ar, out=torch.empty_like(Lb.y))
        dx = E.bn_bwd(La, du, grads, ar)
        for i in reversed(range(len(self.convs))):
            X, L, sec = s["levels"][i]
            w = _w(self.convs[i].weight)
            da, sums = E.se_bwd(sec, dx, grads, ar, bn_sums=True)
            dy = E.bn_bwd(L, da, grads, ar, sums=sums, conv=self.convs[i])
            gw = E.grad_buf(grads, self.convs[i].weight)
            if gw is not None:
                E.wgrad_conv3x3(X, C, w, dy, C, dims, gw)
            if i > 0 or in_need[0]:
                dx = E.conv(self._taps(Lazy(dy), w, C, transpose=True), C, dims, residual=dx)
        return [dx if in_need[0] else None], grads


# =========================================================================================
class MLFC(_AccxModule):
    """Multi-level feature compilation over a 4-level pyramid (ACC_UNet.py:332-527).

    The 480-channel gather concat is never built: per target level the 1x1 conv is split by
    source level; finer sources are average-pooled first (accx_pool_sum), coarser sources are
    contracted at their own resolution and nearest-upsample-added in the epilogue."""

    def __init__(self, in_filters1, in_filters2, in_filters3, in_filters4, lenn=1, variant="base"):
        super().__init__()
        if variant not in ("base", "w", "lite"):
            raise ValueError(variant)
        self.variant = variant
        self.in_filters1, self.in_filters2 = in_filters1, in_filters2
        self.in_filters3, self.in_filters4 = in_filters3, in_filters4
        self.in_filters = in_filters1 + in_filters2 + in_filters3 + in_filters4
        self.no_param_up = nn.Upsample(scale_factor=2)
        self.no_param_down = nn.AvgPool2d(2)
        for kind in ("cnv_blks", "cnv_mrg", "bns", "bns_mrg"):
            for l in range(1, 5):
                setattr(self, f"{kind}{l}", nn.ModuleList([]))
        if variant == "w":
            self.W = nn.Parameter(torch.zeros(1))
        filters = (in_filters1, in_filters2, in_filters3, in_filters4)
        for i in range(lenn):
            for l, c in enumerate(filters, start=1):
                getattr(self, f"cnv_blks{l}").append(Conv2d_batchnorm(self.in_filters, c, (1, 1)))
                getattr(self, f"cnv_mrg{l}").append(Conv2d_batchnorm(2 * c, c, (1, 1)))
                getattr(self, f"bns{l}").append(nn.BatchNorm2d(c))
                getattr(self, f"bns_mrg{l}").append(nn.BatchNorm2d(c))
        self.act = nn.LeakyReLU()
        self.sqe1 = ChannelSELayer(in_filters1)
        self.sqe2 = ChannelSELayer(in_filters2)
        self.sqe3 = ChannelSELayer(in_filters3)
        self.sqe4 = ChannelSELayer(in_filters4)

    def forward(self, x1, x2, x3, x4):
        return self._run(x1, x2, x3, x4)

    # ---------------------------------------------------------------------------------
    def _fwd(self, xs, training, need):
        ar = Arena(xs[0].device)
        filt = [x.shape[3] for x in xs]
        dims = [tuple(x.shape[:3]) for x in xs]
        for l in range(1, 4):
            if dims[l][1] * (1 << l) != dims[0][1] or dims[l][2] * (1 << l) != dims[0][2]:
                raise ValueError("MLFC expects an exact x2 pyramid")
        if self.variant == "lite":
            outs, ctxs = [], []
            for l in range(4):
                o, c = E.se_fwd(Lazy(xs[l]), getattr(self, f"sqe{l + 1}"), ar, training)
                outs.append(o)
                ctxs.append(c)
            return outs, ({"lite": ctxs, "out_like": _out_like(outs)} if need else None)
        tot = sum(filt)
        offs = [sum(filt[:s]) for s in range(4)]
        mix = E.f32(self.W) if self.variant == "w" else None
        # finer sources average-pooled down to every coarser level
        pooled = {}
        for s in range(3):
            for l in range(s + 1, 4):
                pooled[(s, l)] = E.pool_sum(xs[s], l - s, 1.0 / float(4 ** (l - s)))
        lenn = len(self.cnv_blks1)
        saved_it = None
        Lm = None
        outs, fin = [None] * 4, [None] * 4
        # the four target levels are independent chains (gather conv -> SE -> bns -> merge conv -> SE -> bns_mrg ->
        # final SE): one stream lane and one scratch arena per level
        ars = [Arena(xs[0].device) for _ in range(4)]
        with E.fork_lanes(4) as lanes:
            for i in range(lenn):
                blk, mrg, Lc, Lm = [None] * 4, [None] * 4, [None] * 4, [None] * 4
                for l in range(4):
                    with lanes.lane(l):
                        arl = ars[l]
                        cb = getattr(self, f"cnv_blks{l + 1}")[i]
                        wb = _w(cb.conv1.weight)                                  # [C_l, tot], block order
                        # coarser sources, coarsest first; bf16 storage: each adds the sum so far, so that the
                        # contraction at this level reads one addend (see _hanc_core_fwd)
                        chain = xs[0].dtype == torch.bfloat16
                        adds, r = [], None
                        for s in range(3, l, -1):
                            r = E.conv([Op(Lazy(xs[s]), filt[s], WV(wb, offs[s], tot, 1))], filt[l], dims[s],
                                       out_dtype=E.F32, adds=[(r, 1)] if (chain and r is not None) else ())
                            adds.insert(0, (r, s - l))
                        if chain:
                            adds = adds[:1]
                        ops = [Op(Lazy(pooled[(s, l)] if s < l else xs[l]), filt[s], WV(wb, offs[s], tot, 1))
                               for s in range(l + 1)]
                        stt = arl.take(2 * filt[l]) if training else None
                        t, sv = cb._core_fwd(ops, dims[l], arl, training, adds=adds, stats=stt)
                        Lc[l] = E.bn_lazy(t, stt, getattr(self, f"bns{l + 1}")[i], 2, arl, training)
                        blk[l] = sv
                        cm = getattr(self, f"cnv_mrg{l + 1}")[i]
                        wm = _w(cm.conv1.weight)                                  # [C_l, 2*C_l], K index 2c + j
                        C = filt[l]
                        ops = [Op(Lc[l], C, WV(wm, 0, 2 * C, 2)), Op(Lazy(xs[l]), C, WV(wm, 1, 2 * C, 2))]
                        stt = arl.take(2 * C) if training else None
                        t, sv = cm._core_fwd(ops, dims[l], arl, training, residual=xs[l], mix=mix, stats=stt,
                                             mix_param=self.W if mix is not None else None)
                        Lm[l] = E.bn_lazy(t, stt, getattr(self, f"bns_mrg{l + 1}")[i], 2, arl, training)
                        mrg[l] = sv
                saved_it = (i, blk, Lc, mrg)      # only the last repeat reaches the output (as in the reference)
            for l in range(4):
                with lanes.lane(l):
                    outs[l], fin[l] = E.se_fwd(Lm[l], getattr(self, f"sqe{l + 1}"), ars[l], training)
        saved = None
        if need:
            saved = {"xs": xs, "pooled": pooled, "it": saved_it, "Lm": Lm, "fin": fin, "out_like": _out_like(outs),
                     "mix": mix}
        return outs, saved

    def _bwd(self, s, douts, in_need):
        grads = E.GradPool(self.parameters())
        ar = Arena(douts[0].device)
        if self.variant == "lite":
            return [E.se_bwd(c, d, grads, ar) for c, d in zip(s["lite"], douts)], grads
        xs, pooled, Lm, mix = s["xs"], s["pooled"], s["Lm"], s["mix"]
        i, blk, Lc, mrg = s["it"]
        filt = [x.shape[3] for x in xs]
        dims = [tuple(x.shape[:3]) for x in xs]
        tot = sum(filt)
        offs = [sum(filt[:q]) for q in range(4)]
        dxs = [None] * 4

        def acc(l, g):
            dxs[l] = g if dxs[l] is None else E.add_inplace(dxs[l], g)

        def acc_conv(l, ops, N, dims_, **kw):
            """dxs[l] += contraction: accumulated in the contraction's own epilogue (in place)"""
            if dxs[l] is None:
                dxs[l] = E.conv(ops, N, dims_, **kw)
            else:
                E.conv(ops, N, dims_, out=dxs[l], residual=dxs[l])

        # phase A: the four level chains are independent (lane l only touches dxs[l]) -> parallel lanes
        dys_blk = [None] * 4
        ars = [ar] + [Arena(douts[0].device) for _ in range(3)]
        gmix4 = ar.take(4) if (mix is not None and self.W.requires_grad) else None     # one slot per level (W variant)
        with E.fork_lanes(4) as lanes:
            for l in range(4):
                with lanes.lane(l):
                    arl = ars[l]
                    C = filt[l]
                    cm = getattr(self, f"cnv_mrg{l + 1}")[i]
                    wm = _w(cm.conv1.weight)
                    da, sums = E.se_bwd(s["fin"][l], douts[l], grads, arl, bn_sums=True)      # final SE
                    dt_ = E.bn_bwd(Lm[l], da, grads, arl, sums=sums)        # bns_mrg: grad wrt (SE_out*mix + x*(1-mix))
                    # residual branch
                    if mix is None:
                        acc(l, dt_)          # aliasing is safe: dt_ is only read by the launches queued below
                    else:
                        acc(l, self._scaled(dt_, mix, one_minus=True))
                    dy = cm._core_bwd(mrg[l], dt_, grads, arl,                      # through SE(mix inside) + BN
                                      gmix=None if gmix4 is None else gmix4[l:l + 1])
                    gw = E.grad_buf(grads, cm.conv1.weight)
                    if gw is not None:
                        E.wgrad(Op(Lc[l], C, WV(wm, 0, 2 * C, 2)), dy, C, dims[l], gw)
                        E.wgrad(Op(Lazy(xs[l]), C, WV(wm, 1, 2 * C, 2)), dy, C, dims[l], gw)
                    acc_conv(l, [Op(Lazy(dy), C, WV(wm, 1, 2, 2 * C))], C, dims[l])          # d wrt x (odd K)
                    dac = E.conv([Op(Lazy(dy), C, WV(wm, 0, 2, 2 * C))], C, dims[l])         # d wrt x_c (even K)
                    dtc = E.bn_bwd(Lc[l], dac, grads, arl)                                    # bns
                    cb = getattr(self, f"cnv_blks{l + 1}")[i]
                    dys_blk[l] = cb._core_bwd(blk[l], dtc, grads, arl)
        if gmix4 is not None:
            E.grad_buf(grads, self.W).add_(gmix4.sum())        # the four levels, in level order
        # phase B: gather conv backward; lane `src` owns dxs[src] and collects the contributions of every target level
        gws = [E.grad_buf(grads, getattr(self, f"cnv_blks{l + 1}")[i].conv1.weight) for l in range(4)]
        with E.fork_lanes(4) as lanes:
            for src in range(4):
                with lanes.lane(src):
                    Cs = filt[src]
                    for l in range(4):
                        cb = getattr(self, f"cnv_blks{l + 1}")[i]
                        wb = _w(cb.conv1.weight)
                        gw, dy, C = gws[l], dys_blk[l], filt[l]
                        wv = WV(wb, offs[src], tot, 1)
                        wvt = WV(wb, offs[src], 1, tot)
                        if src <= l:
                            A = Lazy(pooled[(src, l)] if src < l else xs[l])
                            if gw is not None:
                                E.wgrad(Op(A, Cs, wv), dy, C, dims[l], gw)
                            if src == l:
                                acc_conv(l, [Op(Lazy(dy), C, wvt)], Cs, dims[l])
                                continue
                            g = E.conv([Op(Lazy(dy), C, wvt)], Cs, dims[l])
                            first = dxs[src] is None           # gradient of the average pool: broadcast / s^2
                            if first:
                                dxs[src] = torch.empty_like(xs[src])
                            E.upsample_add(g, dxs[src], l - src, 1.0 / float(4 ** (l - src)), accumulate=not first)
                        else:         # coarser source, contracted at its own resolution: block-sum dY first
                            dR = E.pool_sum(dy, src - l, 1.0)
                            if gw is not None:
                                E.wgrad(Op(Lazy(xs[src]), Cs, wv), dR, C, dims[src], gw)
                            acc_conv(src, [Op(Lazy(dR), C, wvt)], Cs, dims[src], out_dtype=E.dt(xs[src]))
        return dxs, grads

    @staticmethod
    def _scaled(t, mix, one_minus):
        # W variant only (tiny extra pass): t * (1 - W) computed on the device without a host sync
        return t * ((1.0 - mix) if one_minus else mix).to(t.dtype)
