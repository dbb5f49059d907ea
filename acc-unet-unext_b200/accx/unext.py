"""UNeXt's shifted tokenized-MLP block on the accx kernels (BASELINE.json configs[2]).

Drop-ins for /root/reference/Experiments/nets/UNext.py -- same class names, constructor signatures, submodule tree
(state_dict keys, initialiser RNG order) and forward signatures:

    DWConv(dim)                                                   UNext.py:150-160
    shiftmlp(in_features, hidden_features, out_features, ...)     UNext.py:38-113
    shiftedBlock(dim, num_heads, mlp_ratio, ..., norm_layer)      UNext.py:117-147
    OverlapPatchEmbed, UNext                                      UNext.py:162-358  (the measurement vehicle: its conv
                                                                  stem / decoder stay torch operators)

Tokens [B, N = H*W, C] ARE an NHWC tensor.  The shift (pad 2 -> chunk(5) over channels -> roll chunk g by g-2 along H
(first) or W (second) -> narrow, UNext.py:78-84,97-103) moves no data here: chunk g of the activation is one operand
of the fc1 / fc2 contraction read at the pixel offset (dy, dx) = (2-g, 0) / (0, 2-g), zero outside the image -- the same
shifted-operand mechanism that implements the 3x3 taps of ResPath.  DWConv is the depthwise 3x3 kernel of HANCBlock
(+bias), GELU and LayerNorm are the two element-wise kernels of csrc/unext_kernels.cu.  Backward is hand-written on
the same kernels.  fp32 or bf16 tokens, fp32 parameters; CUDA only.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import engine as E
from .engine import Arena, Lazy, Op, WV


def _chunks(C: int, n: int = 5):
    """(first channel, width) of torch.chunk(x, n, dim=1) for C channels"""
    size = -(-C // n)
    out, c0 = [], 0
    while c0 < C:
        out.append((c0, min(size, C - c0)))
        c0 += size
    return out


def _colsum(t: torch.Tensor, ar: Arena) -> torch.Tensor:
    """per-channel sum over all pixels of a [.., C] tensor (one accx pass; bias gradients)"""
    C = t.shape[-1]
    st = ar.take(2 * C)
    E.materialize(Lazy(t), stats=st, stats_only=True)
    return st[:C]


def _shift_ops(src: Lazy, w: torch.Tensor, ld: int, pad: int, along_h: bool, n_shift: int):
    """operands of `Linear(shifted(src))`: chunk g is read at offset g - pad along H (W); weight [N, ld] row-major"""
    C = src.y.shape[-1]
    ops = []
    for g, (c0, k) in enumerate(_chunks(C, n_shift)):
        s = g - pad                                          # torch.roll(x_c, s): out[i] = x[i - s]
        ops.append(Op(src, k, WV(w, c0, ld, 1), c0, -s if along_h else 0, 0 if along_h else -s))
    return ops


def _shift_linear_bwd(src: Lazy, w: torch.Tensor, dy: torch.Tensor, N: int, dims, pad, along_h, n_shift, gw, need_dx=True):
    """gradients of y = Linear(shifted(src)) given dy [B,H,W,N]: accumulates dW into gw, returns d src"""
    C = src.y.shape[-1]
    dsrc = torch.empty_like(src.y) if need_dx else None
    for g, (c0, k) in enumerate(_chunks(C, n_shift)):
        s = g - pad
        oy, ox = (-s, 0) if along_h else (0, -s)
        if gw is not None:
            E.wgrad(Op(src, k, WV(w, c0, C, 1), c0, oy, ox), dy, N, dims, gw)
        if need_dx:      # d src[p, c0 + j] = sum_n dy[p - offset, n] * w[n, c0 + j]
            E.conv([Op(Lazy(dy), N, WV(w, c0, 1, C), 0, -oy, -ox)], k, dims, out=dsrc, out_coff=c0)
    return dsrc


class _ShiftMlpFn(torch.autograd.Function):
    """(optional LayerNorm) -> shift_H -> fc1 -> dw3x3 + bias -> GELU -> shift_W -> fc2 (-> + x) on tokens [B, N, C]"""

    @staticmethod
    def forward(ctx, x, H, W, mlp, norm, residual, *params):
        E.require_cuda(x)
        B, N, C = x.shape
        if N != H * W:
            raise ValueError(f"shiftmlp: {N} tokens do not form a {H} x {W} map")
        hidden, Cout = mlp.fc1.out_features, mlp.fc2.out_features
        if hidden != C:
            raise ValueError("shiftmlp: hidden_features must equal in_features (the reference reshapes the hidden tokens "
                             "with the INPUT channel count, UNext.py:94)")
        if residual and Cout != C:
            raise ValueError("shiftedBlock: out_features must equal dim")
        xn = x.detach().contiguous().view(B, H, W, C)
        dims = (B, H, W)
        ar = Arena(x.device)
        mean = rstd = None
        y0 = xn
        if norm is not None:
            y0 = torch.empty_like(xn)
            mean, rstd = ar.take(B * N), ar.take(B * N)
            E.layernorm_fwd(xn, E.f32(norm.weight), E.f32(norm.bias), norm.eps, y0, mean, rstd)
        w1, w2, wd = E.f32(mlp.fc1.weight), E.f32(mlp.fc2.weight), E.f32(mlp.dwconv.dwconv.weight)
        L0 = Lazy(y0)
        y1 = E.conv(_shift_ops(L0, w1, C, mlp.pad, True, mlp.shift_size), hidden, dims, bias=E.f32(mlp.fc1.bias))
        y2 = E.dw_fwd(Lazy(y1), wd, E.f32(mlp.dwconv.dwconv.bias), None)
        a = E.gelu(y2)
        y3 = E.conv(_shift_ops(Lazy(a), w2, hidden, mlp.pad, False, mlp.shift_size), Cout, dims, bias=E.f32(mlp.fc2.bias),
                    residual=xn if residual else None)
        ctx.saved = (xn, y0, mean, rstd, y1, y2, a)
        ctx.mods = (mlp, norm, residual, dims)
        ctx.params = params
        return y3.view(B, N, Cout)

    @staticmethod
    def backward(ctx, dout):
        xn, y0, mean, rstd, y1, y2, a = ctx.saved
        ctx.saved = None
        mlp, norm, residual, dims = ctx.mods
        B, H, W = dims
        C, hidden, Cout = xn.shape[-1], mlp.fc1.out_features, mlp.fc2.out_features
        d3 = dout.detach()
        if d3.dtype != xn.dtype:
            d3 = d3.to(xn.dtype)
        d3 = d3.contiguous().view(B, H, W, Cout)
        ar = Arena(xn.device)
        grads = E.GradPool(ctx.params)
        w1, w2, wd = E.f32(mlp.fc1.weight), E.f32(mlp.fc2.weight), E.f32(mlp.dwconv.dwconv.weight)
        E.BWD_DEPTH[0] += 1
        try:
            gb = E.grad_buf(grads, mlp.fc2.bias)
            if gb is not None:
                gb.add_(_colsum(d3, ar))
            da = _shift_linear_bwd(Lazy(a), w2, d3, Cout, dims, mlp.pad, False, mlp.shift_size, E.grad_buf(grads, mlp.fc2.weight))
            d2 = E.gelu_bwd(y2, da)
            gb = E.grad_buf(grads, mlp.dwconv.dwconv.bias)
            if gb is not None:
                gb.add_(_colsum(d2, ar))
            gwd = E.grad_buf(grads, mlp.dwconv.dwconv.weight)
            if gwd is not None:
                E.dw_wgrad(Lazy(y1), d2, gwd)
            d1 = E.dw_fwd(Lazy(d2), wd, None, None, flip=True)
            gb = E.grad_buf(grads, mlp.fc1.bias)
            if gb is not None:
                gb.add_(_colsum(d1, ar))
            need_dx = ctx.needs_input_grad[0]
            d0 = _shift_linear_bwd(Lazy(y0), w1, d1, hidden, dims, mlp.pad, True, mlp.shift_size,
                                   E.grad_buf(grads, mlp.fc1.weight), need_dx=need_dx or norm is not None)
            dx = d0
            if norm is not None:
                dx = torch.empty_like(xn)
                E.layernorm_bwd(xn, E.f32(norm.weight), mean, rstd, d0, dx, E.grad_buf(grads, norm.weight),
                                E.grad_buf(grads, norm.bias))
            if residual and dx is not None:
                E.add_inplace(dx, d3)
        finally:
            E.BWD_DEPTH[0] -= 1
            E.module_backward_end()
        gp = E.param_grads(ctx.params, grads)
        return (dx.view(B, H * W, C) if need_dx else None, None, None, None, None, None, *gp)


class DWConv(nn.Module):
    """depthwise 3x3 + bias on tokens (UNext.py:150-160)"""

    def __init__(self, dim=768):
        super().__init__()
        self.dwconv = nn.Conv2d(dim, dim, 3, 1, 1, bias=True, groups=dim)

    def forward(self, x, H, W):
        return _DWConvFn.apply(x, H, W, self.dwconv.weight, self.dwconv.bias)


class _DWConvFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, H, W, weight, bias):
        E.require_cuda(x)
        B, N, C = x.shape
        xn = x.detach().contiguous().view(B, H, W, C)
        y = E.dw_fwd(Lazy(xn), E.f32(weight), E.f32(bias), None)
        ctx.xn, ctx.weight, ctx.bias = xn, weight, bias
        return y.view(B, N, C)

    @staticmethod
    def backward(ctx, dy):
        xn, weight, bias = ctx.xn, ctx.weight, ctx.bias
        ctx.xn = None
        B, H, W, C = xn.shape
        d = dy.detach().to(xn.dtype).contiguous().view(B, H, W, C)
        grads = E.GradPool([weight, bias])
        ar = Arena(xn.device)
        if ctx.needs_input_grad[4]:
            E.grad_buf(grads, bias).add_(_colsum(d, ar))
        if ctx.needs_input_grad[3]:
            E.dw_wgrad(Lazy(xn), d, E.grad_buf(grads, weight))
        dx = E.dw_fwd(Lazy(d), E.f32(weight), None, None, flip=True).view(B, H * W, C) if ctx.needs_input_grad[0] else None
        E.module_backward_end()
        gw, gb = E.param_grads([weight, bias], grads)
        return dx, None, None, gw, gb


def _init_weights(m):
    """UNext.py:56-69 (timm's trunc_normal_ is torch.nn.init.trunc_normal_: same algorithm, same RNG consumption)"""
    if isinstance(m, nn.Linear):
        nn.init.trunc_normal_(m.weight, std=.02)
        if m.bias is not None:
            nn.init.constant_(m.bias, 0)
    elif isinstance(m, nn.LayerNorm):
        nn.init.constant_(m.bias, 0)
        nn.init.constant_(m.weight, 1.0)
    elif isinstance(m, nn.Conv2d):
        fan_out = m.kernel_size[0] * m.kernel_size[1] * m.out_channels
        fan_out //= m.groups
        m.weight.data.normal_(0, math.sqrt(2.0 / fan_out))
        if m.bias is not None:
            m.bias.data.zero_()


class shiftmlp(nn.Module):
    """UNext.py:38-113"""

    def __init__(self, in_features, hidden_features=None, out_features=None, act_layer=nn.GELU, drop=0., shift_size=5):
        super().__init__()
        out_features = out_features or in_features
        hidden_features = hidden_features or in_features
        if act_layer is not nn.GELU:
            raise NotImplementedError("accx shiftmlp implements nn.GELU (all the reference builds)")
        if drop != 0.:
            raise NotImplementedError("accx shiftmlp implements drop = 0 (all the reference builds)")
        self.dim = in_features
        self.fc1 = nn.Linear(in_features, hidden_features)
        self.dwconv = DWConv(hidden_features)
        self.act = act_layer()
        self.fc2 = nn.Linear(hidden_features, out_features)
        self.drop = nn.Dropout(drop)
        self.shift_size = shift_size
        self.pad = shift_size // 2
        self.apply(_init_weights)

    def forward(self, x, H, W):
        return _ShiftMlpFn.apply(x, H, W, self, None, False, *self.parameters())


class shiftedBlock(nn.Module):
    """x + shiftmlp(LayerNorm(x))  (UNext.py:117-147)"""

    def __init__(self, dim, num_heads, mlp_ratio=4., qkv_bias=False, qk_scale=None, drop=0., attn_drop=0.,
                 drop_path=0., act_layer=nn.GELU, norm_layer=nn.LayerNorm, sr_ratio=1):
        super().__init__()
        if drop_path > 0.:
            raise NotImplementedError("accx shiftedBlock implements drop_path = 0 (UNext's default)")
        if norm_layer is not nn.LayerNorm:
            raise NotImplementedError("accx shiftedBlock implements nn.LayerNorm")
        self.drop_path = nn.Identity()
        self.norm2 = norm_layer(dim)
        mlp_hidden_dim = int(dim * mlp_ratio)
        self.mlp = shiftmlp(in_features=dim, hidden_features=mlp_hidden_dim, act_layer=act_layer, drop=drop)
        self.apply(_init_weights)

    def forward(self, x, H, W):
        return _ShiftMlpFn.apply(x, H, W, self.mlp, self.norm2, True, *self.parameters())


class OverlapPatchEmbed(nn.Module):
    """UNext.py:162-203 (torch operators: strided dense conv + LayerNorm; outside the hot path)"""

    def __init__(self, img_size=224, patch_size=7, stride=4, in_chans=3, embed_dim=768):
        super().__init__()
        img_size = (img_size, img_size) if isinstance(img_size, int) else tuple(img_size)
        patch_size = (patch_size, patch_size) if isinstance(patch_size, int) else tuple(patch_size)
        self.img_size = img_size
        self.patch_size = patch_size
        self.H, self.W = img_size[0] // patch_size[0], img_size[1] // patch_size[1]
        self.num_patches = self.H * self.W
        self.proj = nn.Conv2d(in_chans, embed_dim, kernel_size=patch_size, stride=stride,
                              padding=(patch_size[0] // 2, patch_size[1] // 2))
        self.norm = nn.LayerNorm(embed_dim)
        self.apply(_init_weights)

    def forward(self, x):
        x = self.proj(x)
        _, _, H, W = x.shape
        x = x.flatten(2).transpose(1, 2)
        x = self.norm(x)
        return x, H, W


class UNext(nn.Module):
    """UNext.py:205-358: conv stem + two tokenized-MLP stages + decoder with two more.  The four shiftedBlocks run the
    accx kernels; the dense 3x3 convs, BatchNorm, pooling and bilinear upsampling of the stem / decoder stay torch
    operators (they are outside the north-star path; this class is the vehicle for BASELINE configs[2])."""

    def __init__(self, n_channels=3, n_classes=1, deep_supervision=False, img_size=224, patch_size=16, in_chans=3,
                 embed_dims=[128, 160, 256], num_heads=[1, 2, 4, 8], mlp_ratios=[4, 4, 4, 4], qkv_bias=False, qk_scale=None,
                 drop_rate=0., attn_drop_rate=0., drop_path_rate=0., norm_layer=nn.LayerNorm, depths=[1, 1, 1],
                 sr_ratios=[8, 4, 2, 1], compute_dtype=None, **kwargs):
        super().__init__()
        # storage dtype of the tokens inside the shiftedBlocks (None: whatever arrives; torch.bfloat16: throughput mode,
        # the same switch ACC_UNet(compute_dtype=) has); the torch stem / decoder keep their own dtype
        self.compute_dtype = compute_dtype
        self.encoder1 = nn.Conv2d(n_channels, 16, 3, stride=1, padding=1)
        self.encoder2 = nn.Conv2d(16, 32, 3, stride=1, padding=1)
        self.encoder3 = nn.Conv2d(32, 128, 3, stride=1, padding=1)
        self.ebn1 = nn.BatchNorm2d(16)
        self.ebn2 = nn.BatchNorm2d(32)
        self.ebn3 = nn.BatchNorm2d(128)
        self.norm3 = norm_layer(embed_dims[1])
        self.norm4 = norm_layer(embed_dims[2])
        self.dnorm3 = norm_layer(160)
        self.dnorm4 = norm_layer(128)
        blk = lambda dim: nn.ModuleList([shiftedBlock(dim=dim, num_heads=num_heads[0], mlp_ratio=1, qkv_bias=qkv_bias,
                                                      qk_scale=qk_scale, drop=drop_rate, attn_drop=attn_drop_rate, drop_path=0.,
                                                      norm_layer=norm_layer, sr_ratio=sr_ratios[0])])
        self.block1 = blk(embed_dims[1])
        self.block2 = blk(embed_dims[2])
        self.dblock1 = blk(embed_dims[1])
        self.dblock2 = blk(embed_dims[0])
        self.patch_embed3 = OverlapPatchEmbed(img_size=img_size // 4, patch_size=3, stride=2, in_chans=embed_dims[0],
                                              embed_dim=embed_dims[1])
        self.patch_embed4 = OverlapPatchEmbed(img_size=img_size // 8, patch_size=3, stride=2, in_chans=embed_dims[1],
                                              embed_dim=embed_dims[2])
        self.decoder1 = nn.Conv2d(256, 160, 3, stride=1, padding=1)
        self.decoder2 = nn.Conv2d(160, 128, 3, stride=1, padding=1)
        self.decoder3 = nn.Conv2d(128, 32, 3, stride=1, padding=1)
        self.decoder4 = nn.Conv2d(32, 16, 3, stride=1, padding=1)
        self.decoder5 = nn.Conv2d(16, 16, 3, stride=1, padding=1)
        self.dbn1 = nn.BatchNorm2d(160)
        self.dbn2 = nn.BatchNorm2d(128)
        self.dbn3 = nn.BatchNorm2d(32)
        self.dbn4 = nn.BatchNorm2d(16)
        self.final = nn.Conv2d(16, n_classes, kernel_size=1)
        self.soft = nn.Softmax(dim=1)

    def _blocks(self, blocks, out, H, W):
        dt0 = out.dtype
        if self.compute_dtype is not None and dt0 != self.compute_dtype:
            out = out.to(self.compute_dtype)
        for blk in blocks:
            out = blk(out, H, W)
        return out if out.dtype == dt0 else out.to(dt0)

    @staticmethod
    def _tokens_to_map(out, B, H, W):
        return out.reshape(B, H, W, -1).permute(0, 3, 1, 2).contiguous()

    def forward(self, x):
        B = x.shape[0]
        up = lambda t: F.interpolate(t, scale_factor=(2, 2), mode="bilinear")
        out = F.relu(F.max_pool2d(self.ebn1(self.encoder1(x)), 2, 2))
        t1 = out
        out = F.relu(F.max_pool2d(self.ebn2(self.encoder2(out)), 2, 2))
        t2 = out
        out = F.relu(F.max_pool2d(self.ebn3(self.encoder3(out)), 2, 2))
        t3 = out
        out, H, W = self.patch_embed3(out)
        out = self._blocks(self.block1, out, H, W)
        out = self._tokens_to_map(self.norm3(out), B, H, W)
        t4 = out
        out, H, W = self.patch_embed4(out)
        out = self._blocks(self.block2, out, H, W)
        out = self._tokens_to_map(self.norm4(out), B, H, W)
        out = F.relu(up(self.dbn1(self.decoder1(out))))
        if t4.shape[2:] != out.shape[2:]:
            t4 = F.interpolate(t4, size=out.shape[2:], mode="bilinear", align_corners=True)
        out = torch.add(out, t4)
        _, _, H, W = out.shape
        out = out.flatten(2).transpose(1, 2)
        out = self._blocks(self.dblock1, out, H, W)
        out = self._tokens_to_map(self.dnorm3(out), B, H, W)
        out = F.relu(up(self.dbn2(self.decoder2(out))))
        if t3.shape[2:] != out.shape[2:]:
            t3 = F.interpolate(t3, size=out.shape[2:], mode="bilinear", align_corners=True)
        out = torch.add(out, t3)
        _, _, H, W = out.shape
        out = out.flatten(2).transpose(1, 2)
        out = self._blocks(self.dblock2, out, H, W)
        out = self._tokens_to_map(self.dnorm4(out), B, H, W)
        out = F.relu(up(self.dbn3(self.decoder3(out))))
        if t2.shape[2:] != out.shape[2:]:
            t2 = F.interpolate(t2, size=out.shape[2:], mode="bilinear", align_corners=True)
        out = torch.add(out, t2)
        out = F.relu(up(self.dbn4(self.decoder4(out))))
        if t1.shape[2:] != out.shape[2:]:
            t1 = F.interpolate(t1, size=out.shape[2:], mode="bilinear", align_corners=True)
        out = torch.add(out, t1)
        out = F.relu(up(self.decoder5(out)))
        out = self.final(out)
        if out.shape[1] == 1:
            out = torch.sigmoid(out)
        return out
